// Exact-fp32 CUDA-core GEMM with generic activation strides.
//
// Used where reduced-precision tensor-core products would break bit-level parity with the reference:
// the weight-normed 1x1 input projection in front of the ResidualVQ search
// (XY_Tokenizer/xy_tokenizer/nn/quantizer.py:224,245 — a TF32 product there would flip code indices),
// and as the in-library cross-check for the tcgen05 kernel in gemm_tc.cu. Operands may be bf16 or
// fp32; products and sums are fp32 FMA in ascending-k order.
#include "common.cuh"
#include "mtts_internal.h"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, kThreads = 256;

struct SimtParams {
  const void* x;
  int rows_per_batch;
  long long xbs, xrs, xks;
  const void* w;
  long long ldw;
  void* out;
  long long ldo;
  int M, N, K;
  int in_bf16, out_bf16, flags;
  const float* bias;
  const float* gamma;
  const void* residual;
  long long ldr;
};

__device__ __forceinline__ float load_in(const void* p, long long idx, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[idx]) : reinterpret_cast<const float*>(p)[idx];
}

__global__ void __launch_bounds__(kThreads) gemm_simt_kernel(const SimtParams p) {
  __shared__ __align__(16) float xs[TK][TM + 4];
  __shared__ __align__(16) float ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool k_fast = (p.xks == 1);
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    // activations tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int ml, kl;
      if (k_fast) {
        ml = tid / 16 + 16 * i;
        kl = tid % 16;
      } else {
        ml = tid % 64;
        kl = tid / 64 + 4 * i;
      }
      const int m = m0 + ml, k = k0 + kl;
      float v = 0.f;
      if (m < p.M && k < p.K) {
        const long long idx = (long long)(m / p.rows_per_batch) * p.xbs + (long long)(m % p.rows_per_batch) * p.xrs +
                              (long long)k * p.xks;
        v = load_in(p.x, idx, p.in_bf16);
      }
      xs[kl][ml] = v;
    }
    // weight tile (always k-contiguous)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nl = tid / 16 + 16 * i, kl = tid % 16;
      const int n = n0 + nl, k = k0 + kl;
      float v = 0.f;
      if (n < p.N && k < p.K) v = load_in(p.w, (long long)n * p.ldw + k, p.in_bf16);
      ws[kl][nl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&ws[k][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.flags & MTTS_EPI_BIAS) v += __ldg(p.bias + n);
      if (p.flags & MTTS_EPI_GELU) v = gelu_erf(v);
      if (p.out_bf16) v = bf16_round(v);
      if (p.flags & MTTS_EPI_GAMMA) v *= __ldg(p.gamma + n);
      if (p.flags & MTTS_EPI_RESIDUAL) {
        const long long ri = (long long)m * p.ldr + n;
        const float r = p.out_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.residual)[ri])
                                   : reinterpret_cast<const float*>(p.residual)[ri];
        v = r + v;
      }
      const long long o = (long long)m * p.ldo + n;
      if (p.out_bf16)
        reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16_rn(v);
      else
        reinterpret_cast<float*>(p.out)[o] = v;
    }
  }
}

}  // namespace

extern "C" int mtts_gemm_simt(const void* x, int rows_per_batch, long long x_batch_stride, long long x_row_stride,
                              long long x_k_stride, const void* w, long long ldw, void* out, long long ldo, int M,
                              int N, int K, int in_dtype, int out_dtype, int flags, const float* bias,
                              const float* gamma, const void* residual, long long ldr, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(M >= 0 && N > 0 && K > 0, "mtts_gemm_simt: bad sizes M=%d N=%d K=%d", M, N, K);
  if (M == 0) return MTTS_OK;
  MTTS_REQUIRE(x && w && out, "mtts_gemm_simt: null pointer");
  MTTS_REQUIRE(rows_per_batch > 0, "mtts_gemm_simt: rows_per_batch must be positive");
  MTTS_REQUIRE(!(flags & MTTS_EPI_SWIGLU), "mtts_gemm_simt: SWIGLU epilogue is tcgen05-only");
  if (flags & MTTS_EPI_BIAS) MTTS_REQUIRE(bias != nullptr, "mtts_gemm_simt: EPI_BIAS without bias");
  if (flags & MTTS_EPI_GAMMA) MTTS_REQUIRE(gamma != nullptr, "mtts_gemm_simt: EPI_GAMMA without gamma");
  if (flags & MTTS_EPI_RESIDUAL) MTTS_REQUIRE(residual != nullptr, "mtts_gemm_simt: EPI_RESIDUAL without residual");
  SimtParams p;
  p.x = x; p.rows_per_batch = rows_per_batch; p.xbs = x_batch_stride; p.xrs = x_row_stride; p.xks = x_k_stride;
  p.w = w; p.ldw = ldw; p.out = out; p.ldo = ldo; p.M = M; p.N = N; p.K = K;
  p.in_bf16 = in_dtype == MTTS_DTYPE_BF16; p.out_bf16 = out_dtype == MTTS_DTYPE_BF16; p.flags = flags;
  p.bias = bias; p.gamma = gamma; p.residual = residual; p.ldr = ldr;
  dim3 grid(ceil_div(N, TN), ceil_div(M, TM));
  gemm_simt_kernel<<<grid, kThreads, 0, stream>>>(p);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
