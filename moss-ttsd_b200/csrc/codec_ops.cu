// fp32 kernels of the XY_Tokenizer decode path that are not GEMMs. Activations are TOKEN-major
// ([batch * frames, channels], channels contiguous) everywhere, so every projection is a plain mtts_gemm and no
// (B,C,T)<->(B,T,C) transposes are ever materialised (the reference transposes around every block,
// XY_Tokenizer/xy_tokenizer/nn/modules.py:1144-1153,1399-1409).
#include <cuda_fp16.h>
#include "common.cuh"
#include "mtts_internal.h"
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel dim (nn.LayerNorm, modules.py:171,182,324,554,1110,1377,1397), optionally
// zeroing rows at or beyond each item's length (`torch.where(attention_mask, hidden_states, 0)`,
// modules.py:407,626). One warp per row.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store4(float* p, float4 o) { *reinterpret_cast<float4*>(p) = o; }
__device__ __forceinline__ void store4(__half* p, float4 o) {  // fp16-operand path of the decoder GEMMs (mtts.h)
  __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, OutT* __restrict__ out,
                                                        long long rows, int C, float eps,
                                                        const int* __restrict__ lengths, int rows_per_item) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  OutT* orow = out + row * C;
  if (lengths) {
    const int item = (int)(row / rows_per_item), t = (int)(row % rows_per_item);
    if (t >= lengths[item]) {
      for (int c = lane * 4; c < C; c += 128) store4(orow + c, make_float4(0.f, 0.f, 0.f, 0.f));
      return;
    }
  }
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
  const float inv = rsqrtf(warp_sum(q) / (float)C + eps);
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    float4 o;
    o.x = (v.x - mean) * inv * ww.x + bb.x;
    o.y = (v.y - mean) * inv * ww.y + bb.y;
    o.z = (v.z - mean) * inv * ww.z + bb.z;
    o.w = (v.w - mean) * inv * ww.w + bb.w;
    store4(orow + c, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Non-causal multi-head attention with per-item key lengths (VarLenAttention.forward, modules.py:117-160).
// qkv: [B*T, 3*E] fp32 (q | k | v, the fused projection incl. biases), head_dim 64. Keys at or beyond an item's
// length are masked (the reference adds finfo.min there and +1.0 — a constant shift — on valid ones,
// modules.py:99-115). Query rows beyond the length are computed over the same valid keys; their values are
// never consumed (they are zeroed after the final LayerNorm, modules.py:407,626).
// Flash-style: CTA = 64 queries x one head, K/V tiles of 64 keys in shared memory, fp32 throughout.
// ------------------------------------------------------------------------------------------------
constexpr int kHD = 64, kQT = 64, kKT = 64, kPitch = 68;

template <bool kExactExp>
__global__ void __launch_bounds__(256) mha_varlen_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                         const int* __restrict__ lengths, int T, int H, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                    // [64][68]
  float* sk = sq + kQT * kPitch;     // [64][68]
  float* sv = sk + kKT * kPitch;     // [64][68]
  float* sp = sv + kKT * kPitch;     // [64][68] probabilities
  const int E = H * kHD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQT;
  const int len = lengths ? min(lengths[b], T) : T;
  const int tid = threadIdx.x;
  const int qi = tid >> 4;   // 0..15 -> queries qi*4 .. qi*4+3
  const int ki = tid & 15;   // 0..15 -> keys ki + 16*j (scores) / dims ki*4..+3 (output)
  const float* base = qkv + (long long)b * T * 3 * E;

  // Q tile (pre-scaled, as `query = q_proj(x) * scaling`, modules.py:126)
  for (int i = tid; i < kQT * (kHD / 4); i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) v = *reinterpret_cast<const float4*>(base + (long long)(q0 + r) * 3 * E + h * kHD + c4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(sq + r * kPitch + c4) = v;
  }
  float m[4], l[4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -1e30f;
    l[i] = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[i][d] = 0.f;
  }
  // all-masked items (len == 0) attend uniformly over every key in the reference; keep that corner defined
  const int kv_len = len > 0 ? len : T;

  for (int k0 = 0; k0 < kv_len; k0 += kKT) {
    __syncthreads();
    for (int i = tid; i < kKT * (kHD / 4); i += 256) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (k0 + r < kv_len) {
        const float* p = base + (long long)(k0 + r) * 3 * E + h * kHD + c4;
        kv4 = *reinterpret_cast<const float4*>(p + E);
        vv4 = *reinterpret_cast<const float4*>(p + 2 * E);
      }
      *reinterpret_cast<float4*>(sk + r * kPitch + c4) = kv4;
      *reinterpret_cast<float4*>(sv + r * kPitch + c4) = vv4;
    }
    __syncthreads();
    // scores: 4 queries x 4 keys per thread
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kHD; d += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(sq + (qi * 4 + i) * kPitch + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(sk + (ki + 16 * j) * kPitch + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
    // online softmax per query row (16 lanes share a row)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + ki + 16 * j >= kv_len) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      const float mn = fmaxf(m[i], mx);
      const float corr = kExactExp ? expf(m[i] - mn) : __expf(m[i] - mn);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pr = kExactExp ? expf(s[i][j] - mn) : __expf(s[i][j] - mn);
        ps += pr;
        sp[(qi * 4 + i) * kPitch + ki + 16 * j] = pr;
      }
      ps += __shfl_xor_sync(0xffffffffu, ps, 8);
      ps += __shfl_xor_sync(0xffffffffu, ps, 4);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      m[i] = mn;
      l[i] = l[i] * corr + ps;
#pragma unroll
      for (int d = 0; d < 4; ++d) acc[i][d] *= corr;
    }
    __syncwarp();  // a query row's probabilities are written and read by the same half-warp
    // O += P V : 4 queries x 4 dims per thread
#pragma unroll 8
    for (int k = 0; k < kKT; ++k) {
      const float4 vv = *reinterpret_cast<const float4*>(sv + k * kPitch + ki * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pr = sp[(qi * 4 + i) * kPitch + k];
        acc[i][0] = fmaf(pr, vv.x, acc[i][0]);
        acc[i][1] = fmaf(pr, vv.y, acc[i][1]);
        acc[i][2] = fmaf(pr, vv.z, acc[i][2]);
        acc[i][3] = fmaf(pr, vv.w, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + qi * 4 + i;
    if (q < T) {
      const float inv = 1.0f / l[i];
      *reinterpret_cast<float4*>(out + ((long long)b * T + q) * E + h * kHD + ki * 4) =
          make_float4(acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core version of the same attention (TF32 mma.sync m16n8k8, fp32 accumulate, flash-style online softmax).
// CTA = 64 queries x one head, 4 warps x 16 query rows; K/V tiles of 64 keys staged in shared memory as TF32.
// S = Q K^T uses Q (A, row-major) and K (B, "col-major" = key-major rows) fragments read straight from the padded
// tiles (pitch 68 words -> conflict-free). For O += P V the probabilities stay in the accumulator layout: thread
// (g, t) holds P[g][8j+2t] and P[g][8j+2t+1]; feeding them as A-fragment elements k = t and k = t+4 simply permutes
// the keys of the block, and the V fragment is read with the same permutation (rows 8j+2t and 8j+2t+1), so no
// register shuffles are needed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) mha_varlen_tc_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                            const int* __restrict__ lengths, int T, int H,
                                                            float scale_log2) {
  extern __shared__ __align__(16) uint32_t smu[];
  uint32_t* sk = smu;                   // [64][68] tf32
  uint32_t* sv = sk + kKT * kPitch;     // [64][68] tf32
  const int E = H * kHD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQT;
  const int len = lengths ? min(lengths[b], T) : T;
  const int kv_len = len > 0 ? len : T;  // all-masked item: uniform over every key, as in the reference
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * T * 3 * E;

  // Q fragments for this warp's 16 rows, pre-scaled by head_dim^-0.5 * log2(e): 8 k-blocks x 4 regs
  uint32_t qa[8][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const float* p0 = base + (long long)min(r0, T - 1) * 3 * E + h * kHD;
    const float* p1 = base + (long long)min(r1, T - 1) * 3 * E + h * kHD;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      qa[kk][0] = to_tf32(p0[kk * 8 + t] * scale_log2);
      qa[kk][1] = to_tf32(p1[kk * 8 + t] * scale_log2);
      qa[kk][2] = to_tf32(p0[kk * 8 + t + 4] * scale_log2);
      qa[kk][3] = to_tf32(p1[kk * 8 + t + 4] * scale_log2);
    }
  }
  float o[8][4];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nb][i] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;  // rows g and g+8

  for (int k0 = 0; k0 < kv_len; k0 += kKT) {
    __syncthreads();
    for (int i = tid; i < kKT * (kHD / 4); i += 128) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (k0 + r < kv_len) {
        const float* p = base + (long long)(k0 + r) * 3 * E + h * kHD + c4;
        kv4 = *reinterpret_cast<const float4*>(p + E);
        vv4 = *reinterpret_cast<const float4*>(p + 2 * E);
      }
      *reinterpret_cast<uint4*>(sk + r * kPitch + c4) = make_uint4(to_tf32(kv4.x), to_tf32(kv4.y), to_tf32(kv4.z), to_tf32(kv4.w));
      *reinterpret_cast<uint4*>(sv + r * kPitch + c4) = make_uint4(to_tf32(vv4.x), to_tf32(vv4.y), to_tf32(vv4.z), to_tf32(vv4.w));
    }
    __syncthreads();
    // ---- S = Q K^T : 8 key blocks of 8
    float sc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[nb][i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t b0 = sk[(nb * 8 + g) * kPitch + kk * 8 + t];
        const uint32_t b1 = sk[(nb * 8 + g) * kPitch + kk * 8 + t + 4];
        mma_tf32_16x8x8(sc[nb], qa[kk], b0, b1);
      }
    }
    // ---- mask keys beyond the item's length, online softmax for rows g (c0,c1) and g+8 (c2,c3)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int key = k0 + nb * 8 + 2 * t;
      if (key >= kv_len) { sc[nb][0] = -INFINITY; sc[nb][2] = -INFINITY; }
      if (key + 1 >= kv_len) { sc[nb][1] = -INFINITY; sc[nb][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(sc[nb][0], sc[nb][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nb][2], sc[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      sc[nb][0] = exp2f(sc[nb][0] - mx0);
      sc[nb][1] = exp2f(sc[nb][1] - mx0);
      sc[nb][2] = exp2f(sc[nb][2] - mx1);
      sc[nb][3] = exp2f(sc[nb][3] - mx1);
      ps0 += sc[nb][0] + sc[nb][1];
      ps1 += sc[nb][2] + sc[nb][3];
    }
    l0 = l0 * c0 + ps0;  // per-thread partial row sums; the quad is reduced once at the end
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
    }
    // ---- O += P V : key blocks j of 8 (k of the MMA), d blocks nb of 8 (n of the MMA)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t pa[4];
      pa[0] = to_tf32(sc[j][0]);  // row g,   key 8j+2t   -> k = t
      pa[1] = to_tf32(sc[j][2]);  // row g+8, key 8j+2t   -> k = t
      pa[2] = to_tf32(sc[j][1]);  // row g,   key 8j+2t+1 -> k = t+4
      pa[3] = to_tf32(sc[j][3]);  // row g+8, key 8j+2t+1 -> k = t+4
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const uint32_t b0 = sv[(j * 8 + 2 * t) * kPitch + nb * 8 + g];
        const uint32_t b1 = sv[(j * 8 + 2 * t + 1) * kPitch + nb * 8 + g];
        mma_tf32_16x8x8(o[nb], pa, b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    if (r0 < T)
      *reinterpret_cast<float2*>(out + ((long long)b * T + r0) * E + h * kHD + nb * 8 + 2 * t) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
    if (r1 < T)
      *reinterpret_cast<float2*>(out + ((long long)b * T + r1) * E + h * kHD + nb * 8 + 2 * t) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------
// fp16-operand version of the tensor-core attention (mma.sync m16n8k16, fp32 accumulate / softmax): Q, K, V and the
// probabilities carry the same 10-bit mantissa as the TF32 kernel above at twice the MMA rate and half the shared-memory
// traffic. CTA = 64 queries x one head, 4 warps x 16 query rows; K/V tiles of 64 keys staged as fp16 [key][dim], pitch
// 72 halfs (conflict-free for the 32-bit K fragment loads and for ldmatrix). S = Q K^T reads K fragments directly;
// O += P V takes P from the score accumulators (two adjacent 8-key blocks form one 16-key A fragment, no shuffles) and
// V through ldmatrix.trans.
// ------------------------------------------------------------------------------------------------
constexpr int kHPitch = 72;  // halfs per staged row

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(128) mha_varlen_h_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                           const int* __restrict__ lengths, int T, int H,
                                                           float scale_log2) {
  __shared__ __align__(16) __half sk[kKT * kHPitch];
  __shared__ __align__(16) __half sv[kKT * kHPitch];
  const int E = H * kHD;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQT;
  const int len = lengths ? min(lengths[b], T) : T;
  const int kv_len = len > 0 ? len : T;  // all-masked item: uniform over every key, as in the reference
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * T * 3 * E;

  uint32_t qa[4][4];  // 4 k-steps of 16 dims
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const float* p0 = base + (long long)min(r0, T - 1) * 3 * E + h * kHD;
    const float* p1 = base + (long long)min(r1, T - 1) * 3 * E + h * kHD;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float2 a0 = *reinterpret_cast<const float2*>(p0 + ks * 16 + 2 * t);
      const float2 a1 = *reinterpret_cast<const float2*>(p1 + ks * 16 + 2 * t);
      const float2 a2 = *reinterpret_cast<const float2*>(p0 + ks * 16 + 2 * t + 8);
      const float2 a3 = *reinterpret_cast<const float2*>(p1 + ks * 16 + 2 * t + 8);
      qa[ks][0] = pack_h2(a0.x * scale_log2, a0.y * scale_log2);
      qa[ks][1] = pack_h2(a1.x * scale_log2, a1.y * scale_log2);
      qa[ks][2] = pack_h2(a2.x * scale_log2, a2.y * scale_log2);
      qa[ks][3] = pack_h2(a3.x * scale_log2, a3.y * scale_log2);
    }
  }
  float o[8][4];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nb][i] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
  const uint32_t sv_addr = (uint32_t)__cvta_generic_to_shared(sv);
  // ldmatrix.x4.trans lane address inside a 16-key x 16-dim block: key = (l % 8) + 8 * ((l / 8) % 2), dim = 8 * (l / 16)
  const uint32_t ldm_off = (uint32_t)((((lane & 7) + 8 * ((lane >> 3) & 1)) * kHPitch + 8 * (lane >> 4)) * 2);

  for (int k0 = 0; k0 < kv_len; k0 += kKT) {
    __syncthreads();
    for (int i = tid; i < kKT * (kHD / 4); i += 128) {
      const int r = i >> 4, c4 = (i & 15) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (k0 + r < kv_len) {
        const float* p = base + (long long)(k0 + r) * 3 * E + h * kHD + c4;
        kv4 = *reinterpret_cast<const float4*>(p + E);
        vv4 = *reinterpret_cast<const float4*>(p + 2 * E);
      }
      *reinterpret_cast<uint2*>(sk + r * kHPitch + c4) = make_uint2(pack_h2(kv4.x, kv4.y), pack_h2(kv4.z, kv4.w));
      *reinterpret_cast<uint2*>(sv + r * kHPitch + c4) = make_uint2(pack_h2(vv4.x, vv4.y), pack_h2(vv4.z, vv4.w));
    }
    __syncthreads();
    float sc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[nb][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const __half* kp = sk + (nb * 8 + g) * kHPitch + ks * 16 + 2 * t;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kp + 8);
        mma_f16_16x8x16(sc[nb], qa[ks], b0, b1);
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int key = k0 + nb * 8 + 2 * t;
      if (key >= kv_len) { sc[nb][0] = -INFINITY; sc[nb][2] = -INFINITY; }
      if (key + 1 >= kv_len) { sc[nb][1] = -INFINITY; sc[nb][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(sc[nb][0], sc[nb][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nb][2], sc[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      sc[nb][0] = exp2f(sc[nb][0] - mx0);
      sc[nb][1] = exp2f(sc[nb][1] - mx0);
      sc[nb][2] = exp2f(sc[nb][2] - mx1);
      sc[nb][3] = exp2f(sc[nb][3] - mx1);
      ps0 += sc[nb][0] + sc[nb][1];
      ps1 += sc[nb][2] + sc[nb][3];
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16-key blocks
      uint32_t pa[4];
      pa[0] = pack_h2(sc[2 * j][0], sc[2 * j][1]);
      pa[1] = pack_h2(sc[2 * j][2], sc[2 * j][3]);
      pa[2] = pack_h2(sc[2 * j + 1][0], sc[2 * j + 1][1]);
      pa[3] = pack_h2(sc[2 * j + 1][2], sc[2 * j + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-dim blocks
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, sv_addr + (uint32_t)((j * 16 * kHPitch + np * 16) * 2) + ldm_off);
        mma_f16_16x8x16(o[2 * np], pa, vb[0], vb[1]);
        mma_f16_16x8x16(o[2 * np + 1], pa, vb[2], vb[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    if (r0 < T)
      *reinterpret_cast<float2*>(out + ((long long)b * T + r0) * E + h * kHD + nb * 8 + 2 * t) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
    if (r1 < T)
      *reinterpret_cast<float2*>(out + ((long long)b * T + r1) * E + h * kHD + nb * 8 + 2 * t) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------
// ConvNeXt front half: depthwise Conv1d(k=7, pad=3, groups=C) + LayerNorm(C, eps) fused
// (ConvNeXtBlock.forward modules.py:1142-1150). x/out: [B, T, C]; w: [C, 7]; zero padding at each item's ends.
// One CTA (C/4 threads, 4 channels each) walks kDwTokens consecutive tokens of one item with a 7-row sliding window in
// registers: every input row is read once per CTA instead of seven times, the 28 filter taps of a thread stay in
// registers, and each of the two LayerNorm reductions costs one barrier (warp shuffle + per-warp partials).
// ------------------------------------------------------------------------------------------------
constexpr int kDwTokens = 32;

// Two tokens per iteration (one barrier pair serves both, the two LayerNorm reductions of a token pair travel together)
// and the two input rows of the NEXT iteration are requested before this iteration's arithmetic: the first form — one
// token per iteration, its row requested right before use, two barriers per token — ran at 1.9 TB/s (29 % of HBM).
template <typename OutT>
__global__ void dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ cb,
                                  const float* __restrict__ lw, const float* __restrict__ lb, OutT* __restrict__ out,
                                  int T, int C, float eps) {
  __shared__ float red[2][2][2][32];  // [mean | var][parity][token of the pair][warp]
  const int tiles = (T + kDwTokens - 1) / kDwTokens;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kDwTokens;
  const int c = threadIdx.x * 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float4 bias = *reinterpret_cast<const float4*>(cb + c);
  const float4 g = *reinterpret_cast<const float4*>(lw + c);
  const float4 bb = *reinterpret_cast<const float4*>(lb + c);
  float wk[4][7];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) wk[i][j] = __ldg(w + (c + i) * 7 + j);
  const float* xb = x + (long long)b * T * C + c;
  auto row = [&](int tt) {
    return (tt >= 0 && tt < T) ? *reinterpret_cast<const float4*>(xb + (long long)tt * C) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 win[8];  // rows t-3 .. t+4 of the current token pair (t, t+1)
#pragma unroll
  for (int j = 0; j < 6; ++j) win[j + 2] = row(t0 - 3 + j);
  float4 nx0 = row(t0 + 3), nx1 = row(t0 + 4);
  const float invC = 1.0f / (float)C;
  auto conv = [&](int o) {
    float4 a = bias;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      a.x = fmaf(wk[0][j], win[o + j].x, a.x);
      a.y = fmaf(wk[1][j], win[o + j].y, a.y);
      a.z = fmaf(wk[2][j], win[o + j].z, a.z);
      a.w = fmaf(wk[3][j], win[o + j].w, a.w);
    }
    return a;
  };
#pragma unroll 1
  for (int i = 0; i < kDwTokens; i += 2) {
    const int t = t0 + i;
    if (t >= T) break;  // uniform
#pragma unroll
    for (int j = 0; j < 6; ++j) win[j] = win[j + 2];
    win[6] = nx0;
    win[7] = nx1;
    nx0 = row(t + 5);  // rows of the next pair: in flight during this pair's arithmetic and barriers
    nx1 = row(t + 6);
    const bool two = t + 1 < T;  // uniform
    const float4 a0 = conv(0), a1 = conv(1);
    const int par = (i >> 1) & 1;
    float s0 = warp_sum((a0.x + a0.y) + (a0.z + a0.w));
    float s1 = warp_sum((a1.x + a1.y) + (a1.z + a1.w));
    if (lane == 0) { red[0][par][0][warp] = s0; red[0][par][1][warp] = s1; }
    __syncthreads();
    float tot0 = 0.f, tot1 = 0.f;
    for (int k = 0; k < nw; ++k) { tot0 += red[0][par][0][k]; tot1 += red[0][par][1][k]; }
    const float mean0 = tot0 * invC, mean1 = tot1 * invC;
    const float d00 = a0.x - mean0, d01 = a0.y - mean0, d02 = a0.z - mean0, d03 = a0.w - mean0;
    const float d10 = a1.x - mean1, d11 = a1.y - mean1, d12 = a1.z - mean1, d13 = a1.w - mean1;
    s0 = warp_sum((d00 * d00 + d01 * d01) + (d02 * d02 + d03 * d03));
    s1 = warp_sum((d10 * d10 + d11 * d11) + (d12 * d12 + d13 * d13));
    if (lane == 0) { red[1][par][0][warp] = s0; red[1][par][1][warp] = s1; }
    __syncthreads();
    tot0 = 0.f; tot1 = 0.f;
    for (int k = 0; k < nw; ++k) { tot0 += red[1][par][0][k]; tot1 += red[1][par][1][k]; }
    const float inv0 = rsqrtf(tot0 * invC + eps), inv1 = rsqrtf(tot1 * invC + eps);
    store4(out + ((long long)b * T + t) * C + c,
           make_float4(d00 * inv0 * g.x + bb.x, d01 * inv0 * g.y + bb.y, d02 * inv0 * g.z + bb.z, d03 * inv0 * g.w + bb.w));
    if (two)
      store4(out + ((long long)b * T + t + 1) * C + c,
             make_float4(d10 * inv1 * g.x + bb.x, d11 * inv1 * g.y + bb.y, d12 * inv1 * g.z + bb.z, d13 * inv1 * g.w + bb.w));
  }
}

// ------------------------------------------------------------------------------------------------
// ConvTranspose1d as GEMM + gather: y[b, t, j, co] = sum_ci x[b,t,ci] W[ci,co,j] comes from mtts_gemm; this kernel
// overlap-adds the taps:  out[b, u, co] = act(bias[co] + sum_{j : (u-j) % stride == 0, 0 <= (u-j)/stride < Tin} y[b,(u-j)/stride,j,co])
// (OmniAudioDecoder deconv1/deconv2 + GELU + trim, modules.py:354-368,413-419).
// ------------------------------------------------------------------------------------------------
__global__ void convt_gather_kernel(const float* __restrict__ y, const float* __restrict__ bias, float* __restrict__ out,
                                    int B, int Tin, int Cout, int K, int stride, int Tout, int gelu) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = Cout >> 2;
  const long long total = (long long)B * Tout * cv;
  if (gid >= total) return;
  const int c = (int)(gid % cv) * 4;
  const int u = (int)((gid / cv) % Tout);
  const int b = (int)(gid / ((long long)cv * Tout));
  float4 a = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < K; ++j) {
    const int r = u - j;
    if (r < 0 || (r % stride) != 0) continue;
    const int t = r / stride;
    if (t >= Tin) continue;
    const float4 v = *reinterpret_cast<const float4*>(y + (((long long)b * Tin + t) * K + j) * Cout + c);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  if (gelu) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
  *reinterpret_cast<float4*>(out + ((long long)b * Tout + u) * Cout + c) = a;
}

// im2col for Conv1d(k, pad=(k-1)/2) on token-major input: col[b,t, j*Cin + ci] = x[b, t + j - pad, ci] (0 outside).
// (VocosBackbone.embed, modules.py:1372,1401.)
__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int T, int Cin, int K,
                              int ld_col, int stride, int Tout) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = Cin >> 2;
  const long long total = (long long)B * Tout * K * cv;
  if (gid >= total) return;
  const int c = (int)(gid % cv) * 4;
  const int j = (int)((gid / cv) % K);
  const long long otok = gid / ((long long)cv * K);
  const int to = (int)(otok % Tout);
  const long long b = otok / Tout;
  const int tt = to * stride + j - (K - 1) / 2;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tt >= 0 && tt < T) v = *reinterpret_cast<const float4*>(x + (b * T + tt) * Cin + c);
  *reinterpret_cast<float4*>(col + otok * ld_col + j * Cin + c) = v;
}

// ------------------------------------------------------------------------------------------------
// Log-mel front end (MelFeatureExtractor._torch_extract_fbank_features, nn/feature_extractor.py:78-104):
// torch.stft(n_fft, hop, hann, center=True -> reflect pad n_fft/2) restated as framing + an exact-fp32 DFT GEMM.
//   stft_frames: frames[b, t, n] = window[n] * xpad[b, t*hop + n - n_fft/2]  (reflect at both ends of the L-sample
//                chunk, which the extractor has already zero-padded to 30 s), t < T.
//   power:       P[r, k] = re^2 + im^2 from the GEMM output [r, (re_0..re_{F-1} | im_0..im_{F-1})]
//   logmel_finish (one CTA per item): log10(max(mel, 1e-10)); clamp at (item max - 8); (x + 4) / 4.
// ------------------------------------------------------------------------------------------------
__global__ void stft_frames_kernel(const float* __restrict__ wav, long long ld_wav, const float* __restrict__ window,
                                   float* __restrict__ frames, int B, int T, int L, int n_fft, int hop) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * n_fft;
  if (gid >= total) return;
  const int n = (int)(gid % n_fft);
  const int t = (int)((gid / n_fft) % T);
  const long long b = gid / ((long long)n_fft * T);
  long long i = (long long)t * hop + n - n_fft / 2;
  if (i < 0) i = -i;
  if (i >= L) i = 2LL * (L - 1) - i;
  frames[gid] = window[n] * wav[b * ld_wav + i];
}

__global__ void power_kernel(const float* __restrict__ spec, long long lds, float* __restrict__ out, long long ldo,
                             long long rows, int F) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * ldo) return;
  const long long r = gid / ldo;
  const int k = (int)(gid % ldo);
  float v = 0.f;
  if (k < F) {
    const float re = spec[r * lds + k], im = spec[r * lds + F + k];
    v = re * re + im * im;
  }
  out[gid] = v;
}

__global__ void __launch_bounds__(1024) logmel_finish_kernel(float* __restrict__ mel, int per_item) {
  __shared__ float red[33];
  float* m = mel + (long long)blockIdx.x * per_item;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < per_item; i += blockDim.x) {
    const float v = log10f(fmaxf(m[i], 1e-10f));
    m[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  for (int i = threadIdx.x; i < per_item; i += blockDim.x) m[i] = (fmaxf(m[i], mx - 8.0f) + 4.0f) / 4.0f;
}

// ------------------------------------------------------------------------------------------------
// ISTFT head (ISTFTHead.forward modules.py:969-988): x = Linear(h) -> (mag, phase); mag = min(exp(mag), 100);
// S = mag * (cos p + i sin p). Writes [Re_0..Re_{F-1}, Im_0..Im_{F-1}, 0 pad] per frame, which the next GEMM
// multiplies with the windowed inverse-DFT basis.
// ------------------------------------------------------------------------------------------------
__global__ void istft_spec_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ spec, long long lds,
                                  long long rows, int F) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * F) return;
  const long long r = gid / F;
  const int k = (int)(gid % F);
  const float mg = fminf(expf(x[r * ldx + k]), 100.0f);
  const float ph = x[r * ldx + F + k];
  spec[r * lds + k] = mg * cosf(ph);
  spec[r * lds + F + k] = mg * sinf(ph);
  if (k == 0)
    for (long long z = 2LL * F; z < lds; ++z) spec[r * lds + z] = 0.f;
}

// Overlap-add of windowed frames + envelope normalisation + "same" trim (ISTFT.forward modules.py:759-792):
// frames [B, T, n_fft] already multiplied by the window; out [B, T*hop].
__global__ void istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                 float* __restrict__ out, int B, int T, int n_fft, int hop) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)T * hop;
  if (gid >= (long long)B * per) return;
  const int b = (int)(gid / per);
  const long long so = gid % per;
  const int pad = (n_fft - hop) / 2;
  const long long s = so + pad;  // position in the untrimmed signal
  int f_hi = (int)(s / hop);
  if (f_hi > T - 1) f_hi = T - 1;
  float acc = 0.f, env = 0.f;
  for (int f = f_hi; f >= 0; --f) {
    const long long o = s - (long long)f * hop;
    if (o >= n_fft) break;
    acc += frames[((long long)b * T + f) * n_fft + o];
    const float wv = window[o];
    env = fmaf(wv, wv, env);
  }
  out[gid] = acc / env;
}

// x[r, :] += table[(r % mod), :]   (sinusoid positional embedding add, modules.py:398-402,600-606)
__global__ void add_rows_mod_kernel(float* __restrict__ x, const float* __restrict__ table, long long rows, int C,
                                    int mod) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C >> 2;
  if (gid >= rows * cv) return;
  const long long r = gid / cv;
  const int c = (int)(gid % cv) * 4;
  float4 v = *reinterpret_cast<float4*>(x + r * C + c);
  const float4 t = *reinterpret_cast<const float4*>(table + (r % mod) * C + c);
  v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  *reinterpret_cast<float4*>(x + r * C + c) = v;
}

// Chunk bookkeeping of XY_Tokenizer.encode / decode (model.py:214-216,241-243: one Python slice copy per item per
// window): dst[b, i] = i < lens[b] ? src[b, i] : 0 for all items of a window in one launch.
template <typename T>
__global__ void rows_prefix_copy_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd,
                                        const int* __restrict__ lens, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[(long long)b * ldd + i] = i < lens[b] ? src[(long long)b * lds + i] : T(0);
}

// 3xTF32 split (see mtts_split_tf32x3 in mtts.h): hi = x rounded to TF32 (10-bit mantissa, round to nearest), lo = x - hi
// (exact in fp32); the tcgen05 TF32 MMA then truncates lo to its top 10 mantissa bits — a 2^-22 relative error on x.
__global__ void split_tf32x3_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long ldo,
                                    long long rows, int K, int weights_order) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int kv = K >> 2;
  if (gid >= rows * kv) return;
  const long long r = gid / kv;
  const int c = (int)(gid % kv) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
  float4 hi, lo;
  hi.x = __uint_as_float(to_tf32(v.x)); hi.y = __uint_as_float(to_tf32(v.y));
  hi.z = __uint_as_float(to_tf32(v.z)); hi.w = __uint_as_float(to_tf32(v.w));
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
  float* o = out + r * ldo + c;
  *reinterpret_cast<float4*>(o) = hi;
  *reinterpret_cast<float4*>(o + K) = weights_order ? hi : lo;
  *reinterpret_cast<float4*>(o + 2 * K) = weights_order ? lo : hi;
}

}  // namespace

int mtts_configure_codec() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(mha_varlen_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(4 * kQT * kPitch * sizeof(float))));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(mha_varlen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(4 * kQT * kPitch * sizeof(float))));
  return MTTS_OK;
}

extern "C" int mtts_layernorm(const float* x, const float* w, const float* b, float* out, long long rows, int C,
                              float eps, const int* lengths, int rows_per_item, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(C > 0 && C % 4 == 0, "mtts_layernorm: C must be a multiple of 4");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && w && b && out, "mtts_layernorm: null pointer");
  MTTS_REQUIRE(lengths == nullptr || rows_per_item > 0, "mtts_layernorm: rows_per_item must be positive with lengths");
  layernorm_kernel<float><<<(unsigned)ceil_div_ll(rows, 8), 256, 0, stream>>>(x, w, b, out, rows, C, eps, lengths,
                                                                             rows_per_item);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_layernorm_f16(const float* x, const float* w, const float* b, void* out_f16, long long rows, int C,
                                  float eps, const int* lengths, int rows_per_item, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(C > 0 && C % 4 == 0, "mtts_layernorm_f16: C must be a multiple of 4");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && w && b && out_f16, "mtts_layernorm_f16: null pointer");
  MTTS_REQUIRE(lengths == nullptr || rows_per_item > 0, "mtts_layernorm_f16: rows_per_item must be positive with lengths");
  layernorm_kernel<__half><<<(unsigned)ceil_div_ll(rows, 8), 256, 0, stream>>>(x, w, b, reinterpret_cast<__half*>(out_f16),
                                                                              rows, C, eps, lengths, rows_per_item);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_mha_varlen(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads,
                               int head_dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kHD, "mtts_mha_varlen: head_dim must be 64 (got %d)", head_dim);
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(qkv && out, "mtts_mha_varlen: null pointer");
  dim3 grid(ceil_div(T, kQT), num_heads, B);
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("MTTS_MHA_SIMT");
    use_tc = (e && e[0] == '1') ? 0 : 1;
  }
  if (use_tc)
    mha_varlen_tc_kernel<<<grid, 128, 2 * kKT * kPitch * sizeof(uint32_t), stream>>>(
        qkv, out, lengths, T, num_heads, 1.4426950408889634f / sqrtf((float)head_dim));
  else
    mha_varlen_kernel<false><<<grid, 256, 4 * kQT * kPitch * sizeof(float), stream>>>(qkv, out, lengths, T, num_heads,
                                                                                     1.0f / sqrtf((float)head_dim));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_mha_varlen_f16(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads,
                                   int head_dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kHD, "mtts_mha_varlen_f16: head_dim must be 64 (got %d)", head_dim);
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(qkv && out, "mtts_mha_varlen_f16: null pointer");
  dim3 grid(ceil_div(T, kQT), num_heads, B);
  mha_varlen_h_kernel<<<grid, 128, 0, stream>>>(qkv, out, lengths, T, num_heads, 1.4426950408889634f / sqrtf((float)head_dim));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_mha_varlen_fp32(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads,
                                    int head_dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kHD, "mtts_mha_varlen_fp32: head_dim must be 64 (got %d)", head_dim);
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(qkv && out, "mtts_mha_varlen_fp32: null pointer");
  dim3 grid(ceil_div(T, kQT), num_heads, B);
  mha_varlen_kernel<true><<<grid, 256, 4 * kQT * kPitch * sizeof(float), stream>>>(qkv, out, lengths, T, num_heads,
                                                                                    1.0f / sqrtf((float)head_dim));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_rows_prefix_copy(const void* src, long long lds, void* dst, long long ldd, const int* lens, int B, int n,
                                     int elem_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "mtts_rows_prefix_copy: elem_bytes must be 4 or 8");
  if (B <= 0 || n <= 0) return MTTS_OK;
  MTTS_REQUIRE(src && dst && lens, "mtts_rows_prefix_copy: null pointer");
  dim3 grid(ceil_div(n, 256), B);
  if (elem_bytes == 4)
    rows_prefix_copy_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(src), lds,
                                                            reinterpret_cast<float*>(dst), ldd, lens, n);
  else
    rows_prefix_copy_kernel<long long><<<grid, 256, 0, stream>>>(reinterpret_cast<const long long*>(src), lds,
                                                                reinterpret_cast<long long*>(dst), ldd, lens, n);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_split_tf32x3(const float* x, long long ldx, float* out, long long ldo, long long rows, int K,
                                 int weights_order, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(K > 0 && K % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldo >= 3LL * K,
               "mtts_split_tf32x3: K, ldx, ldo must be multiples of 4 and ldo >= 3K");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && out, "mtts_split_tf32x3: null pointer");
  const long long total = rows * (K / 4);
  split_tf32x3_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(x, ldx, out, ldo, rows, K, weights_order);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_dwconv7_ln(const float* x, const float* conv_w, const float* conv_b, const float* ln_w,
                               const float* ln_b, float* out, int B, int T, int C, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(C % 128 == 0 && C <= 4096, "mtts_dwconv7_ln: C must be a multiple of 128 and <= 4096");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && conv_w && conv_b && ln_w && ln_b && out, "mtts_dwconv7_ln: null pointer");
  const int tiles = ceil_div(T, kDwTokens);
  dwconv7_ln_kernel<float><<<(unsigned)((long long)B * tiles), C / 4, 0, stream>>>(x, conv_w, conv_b, ln_w, ln_b, out, T, C, eps);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_dwconv7_ln_f16(const float* x, const float* conv_w, const float* conv_b, const float* ln_w,
                                   const float* ln_b, void* out_f16, int B, int T, int C, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(C % 128 == 0 && C <= 4096, "mtts_dwconv7_ln_f16: C must be a multiple of 128 and <= 4096");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && conv_w && conv_b && ln_w && ln_b && out_f16, "mtts_dwconv7_ln_f16: null pointer");
  const int tiles = ceil_div(T, kDwTokens);
  dwconv7_ln_kernel<__half><<<(unsigned)((long long)B * tiles), C / 4, 0, stream>>>(
      x, conv_w, conv_b, ln_w, ln_b, reinterpret_cast<__half*>(out_f16), T, C, eps);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_convt_gather(const float* y, const float* bias, float* out, int B, int Tin, int Cout, int K,
                                 int stride, int Tout, int gelu, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(Cout % 4 == 0 && K > 0 && stride > 0, "mtts_convt_gather: bad shape");
  if (B <= 0 || Tout <= 0) return MTTS_OK;
  MTTS_REQUIRE(y && out, "mtts_convt_gather: null pointer");
  const long long total = (long long)B * Tout * (Cout / 4);
  convt_gather_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(y, bias, out, B, Tin, Cout, K, stride, Tout,
                                                                            gelu);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_im2col(const float* x, float* col, int B, int T, int Cin, int K, int ld_col, int stride, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(Cin % 4 == 0 && (K & 1) == 1 && ld_col >= K * Cin && ld_col % 4 == 0 && stride >= 1, "mtts_im2col: bad shape");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && col, "mtts_im2col: null pointer");
  const int Tout = (T + 2 * ((K - 1) / 2) - K) / stride + 1;
  const long long total = (long long)B * Tout * K * (Cin / 4);
  im2col_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(x, col, B, T, Cin, K, ld_col, stride, Tout);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_stft_frames(const float* wav, long long ld_wav, const float* window, float* frames, int B, int T, int L,
                                int n_fft, int hop, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(n_fft > 0 && hop > 0 && L > n_fft / 2, "mtts_stft_frames: bad shape");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(wav && window && frames, "mtts_stft_frames: null pointer");
  const long long total = (long long)B * T * n_fft;
  stft_frames_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(wav, ld_wav, window, frames, B, T, L, n_fft, hop);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_power_spectrum(const float* spec, long long lds, float* out, long long ldo, long long rows, int num_bins,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(num_bins > 0 && lds >= 2 * num_bins && ldo >= num_bins, "mtts_power_spectrum: bad shape");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(spec && out, "mtts_power_spectrum: null pointer");
  power_kernel<<<(unsigned)ceil_div_ll(rows * ldo, 256), 256, 0, stream>>>(spec, lds, out, ldo, rows, num_bins);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_logmel_finish(float* mel, int B, int per_item, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (B <= 0 || per_item <= 0) return MTTS_OK;
  MTTS_REQUIRE(mel, "mtts_logmel_finish: null pointer");
  logmel_finish_kernel<<<B, 1024, 0, stream>>>(mel, per_item);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_istft_spec(const float* x, long long ldx, float* spec, long long lds, long long rows, int num_bins,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(num_bins > 0 && ldx >= 2 * num_bins && lds >= 2 * num_bins, "mtts_istft_spec: bad shape");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && spec, "mtts_istft_spec: null pointer");
  istft_spec_kernel<<<(unsigned)ceil_div_ll(rows * num_bins, 256), 256, 0, stream>>>(x, ldx, spec, lds, rows, num_bins);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_istft_ola(const float* frames, const float* window, float* out, int B, int T, int n_fft, int hop,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(n_fft > hop && hop > 0 && (n_fft - hop) % 2 == 0, "mtts_istft_ola: bad n_fft/hop");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(frames && window && out, "mtts_istft_ola: null pointer");
  const long long total = (long long)B * T * hop;
  istft_ola_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(frames, window, out, B, T, n_fft, hop);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

// ISTFTHead.forward as ONE entry point (SURVEY.md 8b `mtts_istft_head`; modules.py:939-988): the head projection
// x . head_w^T + head_b (TF32 tcgen05 GEMM), the (log-mag | phase) -> (Re | Im) map, the inverse DFT with the synthesis
// window folded into `basis` (second GEMM) and the overlap-add / envelope normalisation / "same" trim. The three
// intermediates live in the caller's workspace: [rows, 2F] | [rows, lds] | [rows, n_fft] fp32, each 256-byte aligned.
static inline size_t istft_align(size_t n) { return (n + 255) & ~size_t(255); }

extern "C" size_t mtts_istft_head_workspace_bytes(int B, int T, int n_fft, long long lds) {
  if (B <= 0 || T <= 0 || n_fft <= 0) return 0;
  const size_t rows = (size_t)B * T;
  return istft_align(rows * (n_fft + 2) * 4) + istft_align(rows * (size_t)lds * 4) + istft_align(rows * (size_t)n_fft * 4);
}

extern "C" int mtts_istft_head(const float* x, long long ldx, int channels, const float* head_w, long long ld_head_w,
                               const float* head_b, const float* basis, long long lds, const float* window, float* wav,
                               int B, int T, int n_fft, int hop, void* workspace, size_t workspace_bytes, void* stream_) {
  MTTS_REQUIRE(n_fft > 0 && n_fft % 2 == 0 && channels > 0, "mtts_istft_head: bad n_fft / channels");
  const int F = n_fft / 2 + 1;
  MTTS_REQUIRE(lds >= 2 * F, "mtts_istft_head: lds must hold 2 * (n_fft / 2 + 1) columns");
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && head_w && head_b && basis && window && wav && workspace, "mtts_istft_head: null pointer");
  MTTS_REQUIRE(workspace_bytes >= mtts_istft_head_workspace_bytes(B, T, n_fft, lds) &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "mtts_istft_head: workspace too small or not 256-byte aligned");
  const long long rows = (long long)B * T;
  MTTS_REQUIRE(rows <= 0x7fffffffLL, "mtts_istft_head: too many frames");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* hx = reinterpret_cast<float*>(ws);
  float* spec = reinterpret_cast<float*>(ws + istft_align((size_t)rows * 2 * F * 4));
  float* frames = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(spec) + istft_align((size_t)rows * lds * 4));
  int rc = mtts_gemm(x, ldx, head_w, ld_head_w, hx, 2 * F, (int)rows, 2 * F, channels, MTTS_DTYPE_F32, MTTS_DTYPE_F32,
                     MTTS_EPI_BIAS, head_b, nullptr, nullptr, 0, nullptr, 0, stream_);
  if (rc != MTTS_OK) return rc;
  rc = mtts_istft_spec(hx, 2 * F, spec, lds, rows, F, stream_);  // also zeroes the lds - 2F padding columns
  if (rc != MTTS_OK) return rc;
  rc = mtts_gemm(spec, lds, basis, lds, frames, n_fft, (int)rows, n_fft, (int)lds, MTTS_DTYPE_F32, MTTS_DTYPE_F32, 0, nullptr,
                 nullptr, nullptr, 0, nullptr, 0, stream_);
  if (rc != MTTS_OK) return rc;
  return mtts_istft_ola(frames, window, wav, B, T, n_fft, hop, stream_);
}

extern "C" int mtts_add_rows_mod(float* x, const float* table, long long rows, int C, int mod, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(C % 4 == 0 && mod > 0, "mtts_add_rows_mod: bad shape");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && table, "mtts_add_rows_mod: null pointer");
  add_rows_mod_kernel<<<(unsigned)ceil_div_ll(rows * (C / 4), 256), 256, 0, stream>>>(x, table, rows, C, mod);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
