// ResidualVQ nearest-code search and codebook gather for XY_Tokenizer.
//
// Reference semantics (XY_Tokenizer/xy_tokenizer/nn/quantizer.py):
//   VectorQuantize.forward :167-172   dist = ||e||^2 - (2e) @ C^T + ||c||^2 ; idx = argmax(-dist) (lowest index on ties)
//   VectorQuantize.forward :187       z_q = z_e + (z_q - z_e)            (straight-through, evaluated in fp32)
//   ResidualVQ.forward     :277-327   masked residual, residual -= z_q * mask, quantized_out += z_q * mask
//   ResidualVQ.decode_codes:345-361   emb = sum_i C_i[codes_i]
//
// B200 design: one CTA owns kTV residual vectors for ALL layers (the layers of one vector are strictly
// sequential, different vectors are independent), keeps those residuals in shared memory, and streams the
// 16.8 MB of codebooks from L2 through a cp.async double buffer. All arithmetic is fp32 FMA so that the only
// difference from the reference is the summation order of the 512-deep dot products (near-ties are
// adjudicated in fp64 by the tests). The formula is evaluated with the reference's own rounding points:
// fl(fl(a - M) + b).
#include "common.cuh"
#include "mtts_internal.h"

namespace {

constexpr int kTV = 32;        // residual vectors per CTA
constexpr int kTC = 128;       // codes per tile
constexpr int kKC = 32;        // k elements per staged chunk
constexpr int kCPitch = kKC + 4;  // 36 floats: conflict-free LDS.128 across lanes
constexpr int kThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void rvq_norms_kernel(const float* __restrict__ cb, int rows, int dim, float* __restrict__ norms) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = cb + (long long)row * dim;
  float s = 0.f;
  for (int k = lane; k < dim; k += 32) s = fmaf(p[k], p[k], s);
  s = warp_sum(s);
  if (lane == 0) norms[row] = s;
}

// Issue the cp.async loads of one [kTC codes][kKC k] chunk of a codebook into `dst`.
__device__ __forceinline__ void load_code_chunk(float* dst, const float* __restrict__ cb_layer, int code0, int k0,
                                                int dim) {
  // 128 codes x 32 floats = 1024 float4; 256 threads x 4
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int idx = threadIdx.x + i * kThreads;  // 0..1023
    int code = idx >> 3, k4 = idx & 7;
    cp_async16(dst + code * kCPitch + k4 * 4, cb_layer + (long long)(code0 + code) * dim + k0 + k4 * 4);
  }
}

__global__ void __launch_bounds__(kThreads, 2)
rvq_encode_kernel(const float* __restrict__ z, const uint8_t* __restrict__ valid, const float* __restrict__ codebooks,
                  const float* __restrict__ norms, int N, int nq, int K, int dim, long long* __restrict__ codes,
                  float* __restrict__ zq, float* __restrict__ residual_out) {
  extern __shared__ __align__(16) float smem[];
  const int rpitch = dim + 4;
  float* res = smem;                                // [kTV][rpitch]
  float* ctile = res + kTV * rpitch;                // [2][kTC][kCPitch]
  float* a_sm = ctile + 2 * kTC * kCPitch;          // [kTV] ||e||^2
  float* best_d = a_sm + kTV;                       // [kTV]
  int* best_i = reinterpret_cast<int*>(best_d + kTV);  // [kTV]
  int* valid_sm = best_i + kTV;                     // [kTV]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * kTV;

  // ---- load residual tile (masked rows hold the zero vector: `residual * mask`, quantizer.py:278)
  for (int v = warp; v < kTV; v += kThreads / 32) {
    const int row = row0 + v;
    const bool ok = row < N && (valid == nullptr || valid[row] != 0);
    if (lane == 0) valid_sm[v] = ok ? 1 : 0;
    for (int k = lane * 4; k < dim; k += 128) {
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) val = *reinterpret_cast<const float4*>(z + (long long)row * dim + k);
      *reinterpret_cast<float4*>(res + v * rpitch + k) = val;
    }
  }
  __syncthreads();

  const int vbase = warp * 4;  // this warp's 4 vectors
  const int nchunks = dim / kKC;

  for (int layer = 0; layer < nq; ++layer) {
    const float* cb = codebooks + (long long)layer * K * dim;
    const float* nrm = norms + (long long)layer * K;

    // ---- a = ||e||^2 for this warp's vectors
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* r = res + (vbase + j) * rpitch;
      float s = 0.f;
      for (int k = lane; k < dim; k += 32) s = fmaf(r[k], r[k], s);
      s = warp_sum(s);
      if (lane == 0) {
        a_sm[vbase + j] = s;
        best_d[vbase + j] = INFINITY;
        best_i[vbase + j] = 0;
      }
    }
    __syncwarp();

    const int total = (K / kTC) * nchunks;  // staged chunks this layer
    load_code_chunk(ctile, cb, 0, 0, dim);
    cp_async_commit();

    float acc[4][4];
    for (int s = 0; s < total; ++s) {
      const int tile = s / nchunks, kc = s % nchunks;
      if (s + 1 < total) {
        const int t2 = (s + 1) / nchunks, k2 = (s + 1) % nchunks;
        load_code_chunk(ctile + ((s + 1) & 1) * kTC * kCPitch, cb, t2 * kTC, k2 * kKC, dim);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (kc == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      }
      const float* ct = ctile + (s & 1) * kTC * kCPitch;
      const float* rbase = res + vbase * rpitch + kc * kKC;
#pragma unroll
      for (int k = 0; k < kKC; k += 4) {
        float4 rv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[i] = *reinterpret_cast<const float4*>(rbase + i * rpitch + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const float4*>(ct + (j * 32 + lane) * kCPitch + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j] = fmaf(rv[i].x, cv[j].x, acc[i][j]);
            acc[i][j] = fmaf(rv[i].y, cv[j].y, acc[i][j]);
            acc[i][j] = fmaf(rv[i].z, cv[j].z, acc[i][j]);
            acc[i][j] = fmaf(rv[i].w, cv[j].w, acc[i][j]);
          }
      }
      if (kc == nchunks - 1) {
        // ---- tile epilogue: dist = (a - 2*dot) + ||c||^2, running argmin (lowest index wins ties)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = a_sm[vbase + i];
          float bd = INFINITY;
          int bi = 0x7fffffff;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int code = tile * kTC + j * 32 + lane;
            const float m2 = 2.0f * acc[i][j];
            const float d = __fadd_rn(__fsub_rn(a, m2), __ldg(nrm + code));
            if (d < bd || (d == bd && code < bi)) {
              bd = d;
              bi = code;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) {
              bd = od;
              bi = oi;
            }
          }
          if (lane == 0) {
            const float cur = best_d[vbase + i];
            if (bd < cur || (bd == cur && bi < best_i[vbase + i])) {
              best_d[vbase + i] = bd;
              best_i[vbase + i] = bi;
            }
          }
        }
      }
      __syncthreads();  // everyone is done with buffer (s&1) before it is refilled at s+2
    }

    // ---- commit this layer: codes, straight-through value, residual / zq update
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = vbase + j;
      const int row = row0 + v;
      if (row >= N) continue;
      const int idx = best_i[v];
      if (lane == 0) codes[(long long)layer * N + row] = idx;
      if (!valid_sm[v]) continue;
      const float* c = cb + (long long)idx * dim;
      float* r = res + v * rpitch;
      for (int k = lane * 4; k < dim; k += 128) {
        const float4 e = *reinterpret_cast<const float4*>(r + k);
        const float4 q = __ldg(reinterpret_cast<const float4*>(c + k));
        float4 st;  // z_e + (z_q - z_e)
        st.x = __fadd_rn(e.x, __fsub_rn(q.x, e.x));
        st.y = __fadd_rn(e.y, __fsub_rn(q.y, e.y));
        st.z = __fadd_rn(e.z, __fsub_rn(q.z, e.z));
        st.w = __fadd_rn(e.w, __fsub_rn(q.w, e.w));
        float4 rn = make_float4(__fsub_rn(e.x, st.x), __fsub_rn(e.y, st.y), __fsub_rn(e.z, st.z),
                                __fsub_rn(e.w, st.w));
        *reinterpret_cast<float4*>(r + k) = rn;
        if (zq) {
          float4* zp = reinterpret_cast<float4*>(zq + (long long)row * dim + k);
          float4 acc4 = layer == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *zp;
          acc4.x = __fadd_rn(acc4.x, st.x);
          acc4.y = __fadd_rn(acc4.y, st.y);
          acc4.z = __fadd_rn(acc4.z, st.z);
          acc4.w = __fadd_rn(acc4.w, st.w);
          *zp = acc4;
        }
      }
    }
    __syncwarp();
  }

  // ---- tail: invalid rows keep zq = 0 and their input residual
  for (int j = 0; j < 4; ++j) {
    const int v = vbase + j;
    const int row = row0 + v;
    if (row >= N) continue;
    const bool ok = valid_sm[v] != 0;
    for (int k = lane * 4; k < dim; k += 128) {
      if (zq && (!ok || nq == 0))
        *reinterpret_cast<float4*>(zq + (long long)row * dim + k) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (residual_out) {
        float4 val = ok ? *reinterpret_cast<const float4*>(res + v * rpitch + k)
                        : *reinterpret_cast<const float4*>(z + (long long)row * dim + k);
        *reinterpret_cast<float4*>(residual_out + (long long)row * dim + k) = val;
      }
    }
  }
}

__global__ void rvq_decode_kernel(const long long* __restrict__ codes, long long codes_ld,
                                  const float* __restrict__ codebooks, int N, int nq, int K, int dim,
                                  float* __restrict__ out, int* __restrict__ err_flag) {
  const int vec_per_row = dim >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gid / vec_per_row;
  const int k = static_cast<int>(gid % vec_per_row) * 4;
  if (row >= N) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < nq; ++i) {
    const long long code = codes[(long long)i * codes_ld + row];
    if (code < 0 || code >= K) {
      if (err_flag) *err_flag = 1;
      continue;
    }
    const float4 c = __ldg(reinterpret_cast<const float4*>(codebooks + ((long long)i * K + code) * dim + k));
    acc.x = __fadd_rn(acc.x, c.x);
    acc.y = __fadd_rn(acc.y, c.y);
    acc.z = __fadd_rn(acc.z, c.z);
    acc.w = __fadd_rn(acc.w, c.w);
  }
  *reinterpret_cast<float4*>(out + row * dim + k) = acc;
}

}  // namespace

int mtts_configure_rvq() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  return MTTS_OK;
}

extern "C" int mtts_rvq_codebook_norms(const float* codebooks, int nq, int codebook_size, int dim, float* norms,
                                       void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(codebooks && norms && nq > 0 && codebook_size > 0 && dim > 0, "mtts_rvq_codebook_norms: bad arguments");
  const int rows = nq * codebook_size;
  rvq_norms_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(codebooks, rows, dim, norms);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_rvq_encode(const float* z, const uint8_t* valid, const float* codebooks, const float* norms, int N,
                               int nq, int codebook_size, int dim, long long* codes, float* zq, float* residual_out,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(N >= 0 && nq >= 0, "mtts_rvq_encode: negative sizes");
  if (N == 0) return MTTS_OK;
  MTTS_REQUIRE(z && codebooks && norms && codes, "mtts_rvq_encode: null pointer");
  MTTS_REQUIRE(dim > 0 && dim <= 512 && dim % kKC == 0, "mtts_rvq_encode: dim must be a multiple of %d and <= 512 (got %d)",
               kKC, dim);
  MTTS_REQUIRE(codebook_size > 0 && codebook_size % kTC == 0,
               "mtts_rvq_encode: codebook_size must be a multiple of %d (got %d)", kTC, codebook_size);
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebooks) & 15) == 0,
               "mtts_rvq_encode: z and codebooks must be 16-byte aligned");
  const size_t smem = sizeof(float) * (kTV * (dim + 4) + 2 * kTC * kCPitch + 2 * kTV) + sizeof(int) * 2 * kTV;
  rvq_encode_kernel<<<ceil_div(N, kTV), kThreads, smem, stream>>>(z, valid, codebooks, norms, N, nq, codebook_size,
                                                                  dim, codes, zq, residual_out);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_rvq_decode(const long long* codes, long long codes_ld, const float* codebooks, int N, int nq,
                               int codebook_size, int dim, float* out, int* err_flag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(N >= 0 && nq >= 0, "mtts_rvq_decode: negative sizes");
  if (N == 0) return MTTS_OK;
  MTTS_REQUIRE(codes && codebooks && out, "mtts_rvq_decode: null pointer");
  MTTS_REQUIRE(dim > 0 && dim % 4 == 0, "mtts_rvq_decode: dim must be a multiple of 4");
  const long long total = (long long)N * (dim / 4);
  rvq_decode_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(codes, codes_ld, codebooks, N, nq,
                                                                          codebook_size, dim, out, err_flag);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
