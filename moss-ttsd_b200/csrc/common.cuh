// Shared device/host helpers for libmtts (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmtts is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// Error reporting: every C-ABI entry point returns 0 or a negative code and records a message in a
// thread-local buffer that mtts_last_error() hands back.
// ---------------------------------------------------------------------------------------------
#define MTTS_OK 0
#define MTTS_ERR_ARG (-1)
#define MTTS_ERR_CUDA (-2)
#define MTTS_ERR_UNSUPPORTED (-3)

int mtts_set_error(int code, const char* fmt, ...);

#define MTTS_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return mtts_set_error(MTTS_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define MTTS_CUDA_CHECK(expr)                                                                \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return mtts_set_error(MTTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

void mtts_count_launch();  // bumps the process-wide launch counter (bench.py reports it as gpu_launches)

#define MTTS_LAUNCH_CHECK()                                                                  \
  do {                                                                                       \
    mtts_count_launch();                                                                     \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return mtts_set_error(MTTS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). Every kernel of the decode step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: it may start while its predecessor is still running,
// does whatever does not depend on the predecessor (barrier/TMEM setup, and for the GEMM: streaming its WEIGHTS
// into the shared-memory ring), and only then executes griddepcontrol.wait, which returns once the predecessor
// grid has completed and its writes are visible. Each kernel also fires griddepcontrol.launch_dependents at its
// top so that its own successor can do the same. Kernels must not touch predecessor-dependent memory (reads or
// writes) before pdl_wait().
// ---------------------------------------------------------------------------------------------
bool mtts_pdl_enabled();        // false when MTTS_NO_PDL=1
bool mtts_pdl_small_enabled();  // PDL attribute on the small (non-GEMM) kernels; MTTS_PDL_SMALL=0 turns it off

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t mtts_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                              cudaStream_t stream, int cluster_z, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (mtts_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = cluster_z;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t mtts_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args... args) {
  if (!mtts_pdl_small_enabled()) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  }
  return mtts_launch_cluster(kernel, grid, block, smem, stream, 1, args...);
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int mtts_num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

// ---------------------------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// Round to bf16 and back. cvt.rn.bf16x2.f32 d, a, b puts bf16(a) into the UPPER half of d: with b = +0 the word is the
// fp32 bit pattern of the rounded value — one full-rate F2FP instead of F2F.BF16.F32 (+ a shift), which issues on the
// quarter-rate conversion pipe and was what bounded the prompt-sized q/k-norm + RoPE kernel (65 per thread and row).
// Same round-to-nearest-even result as __float2bfloat16_rn.
__device__ __forceinline__ float bf16_round(float x) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x), "f"(0.0f));
  return __uint_as_float(d);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` must hold 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// 16-byte streaming load that does not pollute L1 (weights / KV are read once per step).
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// two packed bf16 -> two floats (exact: bf16 is the top half of fp32)
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// The same value without the IEEE-division subroutine (its slow-path branch keeps the compiler from interleaving
// independent evaluations): q = x * rcp(y) refined by one Newton step on the residual fma(-q, y, x) — correctly rounded
// like x / y except for rare last-bit cases, which sit 16 bits below the bf16 rounding that follows. y = 1 + e^-x >= 1;
// y = inf (x < -88.7) keeps the unrefined quotient -0, as x / inf gives.
__device__ __forceinline__ float silu_nb(float x) {
  const float y = 1.0f + expf(-x);
  float rc;
  asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(y));
  const float q = x * rc;
  const float q2 = fmaf(fmaf(-q, y, x), rc, q);
  return y < 3.0e38f ? q2 : q;
}
// erf-GELU for the epilogue of the TF32 tensor-core GEMMs (nn.GELU() of the codec, modules.py ConvNeXt / transformer
// MLPs): erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, three orders below the TF32 product error it is applied
// to), evaluated as q = x/2 * erfc(|x|/sqrt2) so that the negative tail keeps its relative accuracy:
//   gelu(x) = x - q (x >= 0),  q (x < 0).        ~14 instructions instead of ~45 for erff().
// GELU for fp16 outputs (the codec's fp16-operand decode path): 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))) with
// minimax-fitted coefficients (|error| <= 2.5e-5 against the erf form over [-8, 8]) and the hardware tanh (MUFU.TANH,
// |error| <= 2^-10.99): the total stays below the rounding of the fp16 value that is stored, at 6 FP32 instructions + 1 MUFU
// per element instead of gelu_fast's 19 + 2. The ConvNeXt pw1 GEMM (K = 512: 2.1 us of MMAs per 128 x 256 tile and SM) was
// bound by the issue slots of its epilogue: 831 instructions per 32 elements, 6.8 us per tile (ncu: 'selected' +
// 'not selected' are the top stall reasons of the epilogue warps).
__device__ __forceinline__ float gelu_tanh5(float x) {
  // the fit holds on [-8, 8]; beyond it tanh is saturated (the odd polynomial itself turns around at |x| = 11.1), so the
  // ARGUMENT is clamped while the factor x / 2 is not: gelu(x) = x or -0 exactly out there
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  float p = fmaf(-3.51516790e-04f, x2, 3.70056460e-02f);
  p = fmaf(p, x2, 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(xc * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// The same on two fp16 values at once (HFMA2 / MUFU.TANH.F16x2: 9 instructions per PAIR): for epilogues whose output is
// fp16 anyway — the polynomial's fp16 rounding moves the result by less than the rounding of the stored value.
__device__ __forceinline__ __half2 gelu_tanh5_h2(__half2 x) {
  const __half2 xc = __hmin2(__hmax2(x, __float2half2_rn(-8.0f)), __float2half2_rn(8.0f));
  const __half2 x2 = __hmul2(xc, xc);
  __half2 p = __hfma2(__float2half2_rn(-3.51516790e-04f), x2, __float2half2_rn(3.70056460e-02f));
  p = __hfma2(p, x2, __float2half2_rn(7.97507884e-01f));
  const __half2 arg = __hmul2(xc, p);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&arg)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  return __hfma2(hx, *reinterpret_cast<const __half2*>(&t), hx);
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float q = 0.5f * x * poly * e;
  return x >= 0.f ? x - q : q;
}

#endif  // __CUDACC__
