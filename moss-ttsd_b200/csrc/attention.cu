// Causal GQA attention over a paged (or contiguous) bf16 KV cache — decode (1 query row per sequence) and
// prefill (tiles of QT consecutive query rows of one sequence) share one kernel.
//
// Replaces HF Qwen3Attention's attention call (installed modeling_qwen3.py:184-219,273-288, selected by
// `attn_implementation` in generation_utils.py:15 / inference.py:29) and DynamicCache's torch.cat growth.
// Left-pad rows of the reference batch are simply not stored: every sequence's cache is compact and RoPE
// positions come from cumsum(mask)-1 (last paragraph), which leaves every real
// token's result unchanged.
//
// The op is HBM/L2-bandwidth bound (2 FMA per byte of K/V): CUDA cores with 16-byte streaming loads, fp32
// online softmax, split-KV across CTAs for small batches with a last-arriver combine (no spin waits).
// Thread layout: 16 lanes cover one 128-wide K/V row (8 dims each); the two half-warps of each of the 4
// warps walk different keys, so one CTA consumes 8 keys per step.
#include "common.cuh"
#include "mtts_internal.h"

namespace {

constexpr int kD = 128;
constexpr int kThreads = 128;
constexpr int kGroups = 8;  // half-warps per CTA


struct AttnParams {
  const bf16* q;    // [rows, Hq, 128]
  const bf16* k_pool;
  const bf16* v_pool;  // [num_pages, Hkv, page_size, 128]
  const int* block_table;
  int max_pages, page_shift;
  const int* tile_row0;   // [tiles] first query row of the tile (NULL: tile index, 1 row per tile)
  const int* tile_nrows;  // [tiles] rows in the tile (<= QT)
  const int* row_seq;     // [rows]  sequence of each row (NULL: row index)
  const int* positions;   // [rows]  position of each query row (it attends keys 0..pos)
  bf16* out;              // [rows, Hq * 128]
  int Hq, Hkv;
  float scale_log2;  // head_dim^-0.5 * log2(e)
  int nsplit;
  float* ws;       // split partials
  int* counters;   // [tiles * Hkv]
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

template <int QT, int G>
__global__ void __launch_bounds__(kThreads) gqa_attn_kernel(const AttnParams p) {
  constexpr int NQ = QT * G;  // query vectors handled per CTA
  constexpr int kUnroll = (QT > 1) ? 4 : 8;  // keys in flight per half-warp (decode: 16 x 16 B loads per lane)
  pdl_launch_dependents();    // the o_proj GEMM may start streaming its weights while attention runs
  pdl_wait();
  extern __shared__ float sm[];
  float* sm_m = sm;                       // [kGroups][NQ]
  float* sm_l = sm_m + kGroups * NQ;      // [kGroups][NQ]
  float* sm_o = sm_l + kGroups * NQ;      // [kGroups][NQ][128]
  __shared__ int s_is_last;

  const int tile = blockIdx.x, hk = blockIdx.y, split = blockIdx.z;
  const int row0 = p.tile_row0 ? p.tile_row0[tile] : tile;
  const int nrows = p.tile_nrows ? p.tile_nrows[tile] : 1;
  const int seq = p.row_seq ? p.row_seq[row0] : row0;
  const int pos0 = p.positions[row0];
  const int kv_len = pos0 + nrows;  // keys 0 .. kv_len-1; row i sees keys <= pos0 + i

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int group = warp * 2 + half;
  const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;

  // split-KV range (multiples of kGroups keys)
  int k_begin = 0, k_end = kv_len;
  if (p.nsplit > 1) {
    const int per = ((kv_len + p.nsplit - 1) / p.nsplit + kGroups - 1) / kGroups * kGroups;
    k_begin = min(kv_len, split * per);
    k_end = min(kv_len, k_begin + per);
  }

  // q fragments (pre-scaled): qf[i][8]
  float qf[NQ][8];
#pragma unroll
  for (int t = 0; t < QT; ++t)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (t < nrows) {
        const uint4 u = *reinterpret_cast<const uint4*>(p.q + ((long long)(row0 + t) * p.Hq + hk * G + g) * kD + l16 * 8);
        unpack8(u, qf[t * G + g]);
#pragma unroll
        for (int d = 0; d < 8; ++d) qf[t * G + g][d] *= p.scale_log2;
      } else {
#pragma unroll
        for (int d = 0; d < 8; ++d) qf[t * G + g][d] = 0.f;
      }
    }

  float m[NQ], l[NQ], acc[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    m[i] = -1e30f;
    l[i] = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[i][d] = 0.f;
  }

  const int page_mask = (1 << p.page_shift) - 1;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)p.Hkv << p.page_shift) * kD;

  for (int kb = k_begin + group; kb < k_end; kb += kGroups * kUnroll) {
    uint4 kv[kUnroll], vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int key = kb + u * kGroups;
      if (key < k_end) {
        const int lp = key >> p.page_shift;
        const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
        const long long off = (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + l16 * 8;
        kv[u] = ld_nc_v4(p.k_pool + off);
        vv[u] = ld_nc_v4(p.v_pool + off);
      } else {
        kv[u] = make_uint4(0, 0, 0, 0);
        vv[u] = make_uint4(0, 0, 0, 0);
      }
    }
    // scores of the batch: s[u][i] (replicated over the 16 lanes of the half-warp after the butterfly)
    float sc[kUnroll][NQ];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float kf[8];
      unpack8(kv[u], kf);
      const int key = kb + u * kGroups;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) s = fmaf(qf[i][d], kf[d], s);
        s += __shfl_xor_sync(hmask, s, 8);
        s += __shfl_xor_sync(hmask, s, 4);
        s += __shfl_xor_sync(hmask, s, 2);
        s += __shfl_xor_sync(hmask, s, 1);
        // beyond the range, or (prefill tiles) beyond this query's causal horizon
        if (key >= k_end || (QT > 1 && key > pos0 + i / G)) s = -INFINITY;
        sc[u][i] = s;
      }
    }
    // one online-softmax update per batch of kUnroll keys
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float mx = m[i];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) mx = fmaxf(mx, sc[u][i]);
      const float corr = exp2f(m[i] - mx);
      m[i] = mx;
      float ps = 0.f;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        sc[u][i] = exp2f(sc[u][i] - mx);
        ps += sc[u][i];
      }
      l[i] = l[i] * corr + ps;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[i][d] *= corr;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float vf[8];
      unpack8(vv[u], vf);
#pragma unroll
      for (int i = 0; i < NQ; ++i)
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[i][d] = fmaf(sc[u][i], vf[d], acc[i][d]);
    }
  }

  // ---- combine the 8 half-warp states of this CTA
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    if (l16 == 0) {
      sm_m[group * NQ + i] = m[i];
      sm_l[group * NQ + i] = l[i];
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) sm_o[(group * NQ + i) * kD + l16 * 8 + d] = acc[i][d];
  }
  __syncthreads();

  const int unit = tile * p.Hkv + hk;
  // each thread finalises one (query vector, dim) pair at a time: NQ * 128 outputs
  float* wsb = p.ws ? p.ws + ((long long)unit * p.nsplit) * NQ * (kD + 2) : nullptr;
  for (int o = threadIdx.x; o < NQ * kD; o += kThreads) {
    const int i = o / kD, d = o % kD;
    float M = -1e30f;
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) M = fmaxf(M, sm_m[gq * NQ + i]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) {
      const float w = exp2f(sm_m[gq * NQ + i] - M);
      L = fmaf(sm_l[gq * NQ + i], w, L);
      O = fmaf(sm_o[(gq * NQ + i) * kD + d], w, O);
    }
    if (p.nsplit == 1) {
      const int t = i / G, g = i % G;
      if (t < nrows)
        p.out[((long long)(row0 + t) * p.Hq + hk * G + g) * kD + d] = __float2bfloat16_rn(O / L);
    } else {
      float* w0 = wsb + (long long)split * NQ * (kD + 2) + i * (kD + 2);
      __stcg(w0 + 2 + d, O);
      if (d == 0) {
        __stcg(w0, M);
        __stcg(w0 + 1, L);
      }
    }
  }
  if (p.nsplit == 1) return;

  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(p.counters + unit, 1);
    const int last = prev == p.nsplit - 1;
    if (last) p.counters[unit] = 0;
    s_is_last = last;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  for (int o = threadIdx.x; o < NQ * kD; o += kThreads) {
    const int i = o / kD, d = o % kD;
    float M = -1e30f;
    for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(wsb + ((long long)s * NQ + i) * (kD + 2)));
    float L = 0.f, O = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float* w0 = wsb + ((long long)s * NQ + i) * (kD + 2);
      const float w = exp2f(__ldcg(w0) - M);
      L = fmaf(__ldcg(w0 + 1), w, L);
      O = fmaf(__ldcg(w0 + 2 + d), w, O);
    }
    const int t = i / G, g = i % G;
    if (t < nrows) p.out[((long long)(row0 + t) * p.Hq + hk * G + g) * kD + d] = __float2bfloat16_rn(O / L);
  }
}

// ------------------------------------------------------------------------------------------------
// Decode specialisation (one query row per sequence). The generic kernel above spends ~190 warp-instructions
// per 512-byte key row because the online-softmax bookkeeping is replicated over the 16 lanes that share a row.
// Here 8 lanes share a row (a warp instruction covers 4 rows), each lane owns dims {8l..8l+7} and {64+8l..64+8l+7}
// so that every LDG.128 of 8 lanes is one full 128-byte line, 4 keys per lane-group are in flight per iteration
// (16 x 16 B per lane), and the softmax state is updated once per batch with ex2.approx.
// ------------------------------------------------------------------------------------------------
constexpr int kDecRows = 16;   // key rows per CTA step: 4 warps x 4 lane-groups
constexpr int kDecU = 4;       // steps per iteration -> 64 keys per CTA iteration

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int G>
__global__ void __launch_bounds__(kThreads) gqa_decode_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* sm_m = sm;                        // [kDecRows][G]
  float* sm_l = sm_m + kDecRows * G;       // [kDecRows][G]
  float* sm_o = sm_l + kDecRows * G;       // [kDecRows][G][128]
  __shared__ int s_is_last;

  const int row = blockIdx.x, hk = blockIdx.y, split = blockIdx.z;
  const int seq = p.row_seq ? p.row_seq[row] : row;
  const int kv_len = p.positions[row] + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 3, l8 = lane & 7;
  const int rg = warp * 4 + grp;

  int k_begin = 0, k_end = kv_len;
  if (p.nsplit > 1) {
    const int per = ((kv_len + p.nsplit - 1) / p.nsplit + kDecRows - 1) / kDecRows * kDecRows;
    k_begin = min(kv_len, split * per);
    k_end = min(kv_len, k_begin + per);
  }

  float qf[G][16];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const bf16* qp = p.q + ((long long)row * p.Hq + hk * G + g) * kD + l8 * 8;
    float lo[8], hi[8];
    unpack8(*reinterpret_cast<const uint4*>(qp), lo);
    unpack8(*reinterpret_cast<const uint4*>(qp + 64), hi);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      qf[g][d] = lo[d] * p.scale_log2;
      qf[g][8 + d] = hi[d] * p.scale_log2;
    }
  }
  float m[G], l[G], acc[G][16];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -1e30f;
    l[g] = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[g][d] = 0.f;
  }
  const int page_mask = (1 << p.page_shift) - 1;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)p.Hkv << p.page_shift) * kD;

  for (int kb = k_begin; kb < k_end; kb += kDecRows * kDecU) {
    uint4 k0[kDecU], k1[kDecU], v0[kDecU], v1[kDecU];
#pragma unroll
    for (int u = 0; u < kDecU; ++u) {
      const int key = min(kb + rg + kDecRows * u, k_end - 1);  // clamp: the load is always legal, the score is masked
      const int lp = key >> p.page_shift;
      const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
      const long long off = (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + l8 * 8;
      k0[u] = ld_nc_v4(p.k_pool + off);
      k1[u] = ld_nc_v4(p.k_pool + off + 64);
      v0[u] = ld_nc_v4(p.v_pool + off);
      v1[u] = ld_nc_v4(p.v_pool + off + 64);
    }
    float sc[kDecU][G];
#pragma unroll
    for (int u = 0; u < kDecU; ++u) {
      float kf[16];
      {
        float a[8], b[8];
        unpack8(k0[u], a);
        unpack8(k1[u], b);
#pragma unroll
        for (int d = 0; d < 8; ++d) { kf[d] = a[d]; kf[8 + d] = b[d]; }
      }
      const bool valid = (kb + rg + kDecRows * u) < k_end;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) s = fmaf(qf[g][d], kf[d], s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        sc[u][g] = valid ? s : -INFINITY;
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float mx = m[g];
#pragma unroll
      for (int u = 0; u < kDecU; ++u) mx = fmaxf(mx, sc[u][g]);
      const float corr = fast_exp2(m[g] - mx);
      m[g] = mx;
      float ps = 0.f;
#pragma unroll
      for (int u = 0; u < kDecU; ++u) {
        sc[u][g] = fast_exp2(sc[u][g] - mx);
        ps += sc[u][g];
      }
      l[g] = l[g] * corr + ps;
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[g][d] *= corr;
    }
#pragma unroll
    for (int u = 0; u < kDecU; ++u) {
      float vf[16];
      {
        float a[8], b[8];
        unpack8(v0[u], a);
        unpack8(v1[u], b);
#pragma unroll
        for (int d = 0; d < 8; ++d) { vf[d] = a[d]; vf[8 + d] = b[d]; }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[g][d] = fmaf(sc[u][g], vf[d], acc[g][d]);
    }
  }

  // ---- combine the 16 lane-group states of this CTA
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (l8 == 0) {
      sm_m[rg * G + g] = m[g];
      sm_l[rg * G + g] = l[g];
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      sm_o[(rg * G + g) * kD + l8 * 8 + d] = acc[g][d];
      sm_o[(rg * G + g) * kD + 64 + l8 * 8 + d] = acc[g][8 + d];
    }
  }
  __syncthreads();
  const int unit = row * p.Hkv + hk;
  float* wsb = p.ws ? p.ws + ((long long)unit * p.nsplit) * G * (kD + 2) : nullptr;
  for (int o = threadIdx.x; o < G * kD; o += kThreads) {
    const int g = o / kD, d = o % kD;
    float M = -1e30f;
#pragma unroll
    for (int r = 0; r < kDecRows; ++r) M = fmaxf(M, sm_m[r * G + g]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int r = 0; r < kDecRows; ++r) {
      const float w = fast_exp2(sm_m[r * G + g] - M);
      L = fmaf(sm_l[r * G + g], w, L);
      O = fmaf(sm_o[(r * G + g) * kD + d], w, O);
    }
    if (p.nsplit == 1) {
      p.out[((long long)row * p.Hq + hk * G + g) * kD + d] = __float2bfloat16_rn(O / L);
    } else {
      float* w0 = wsb + (long long)split * G * (kD + 2) + g * (kD + 2);
      __stcg(w0 + 2 + d, O);
      if (d == 0) {
        __stcg(w0, M);
        __stcg(w0 + 1, L);
      }
    }
  }
  if (p.nsplit == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(p.counters + unit, 1);
    const int last = prev == p.nsplit - 1;
    if (last) p.counters[unit] = 0;
    s_is_last = last;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  for (int o = threadIdx.x; o < G * kD; o += kThreads) {
    const int g = o / kD, d = o % kD;
    float M = -1e30f;
    for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(wsb + ((long long)s * G + g) * (kD + 2)));
    float L = 0.f, O = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float* w0 = wsb + ((long long)s * G + g) * (kD + 2);
      const float w = fast_exp2(__ldcg(w0) - M);
      L = fmaf(__ldcg(w0 + 1), w, L);
      O = fmaf(__ldcg(w0 + 2 + d), w, O);
    }
    p.out[((long long)row * p.Hq + hk * G + g) * kD + d] = __float2bfloat16_rn(O / L);
  }
}

template <int G>
int launch_decode(const AttnParams& p, int rows, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (size_t)kDecRows * G * (kD + 2);
  dim3 grid(rows, p.Hkv, p.nsplit);
  MTTS_CUDA_CHECK(mtts_launch(gqa_decode_kernel<G>, grid, dim3(kThreads), smem, stream, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

template <int QT, int G>
int launch_attn(const AttnParams& p, int tiles, cudaStream_t stream) {
  constexpr int NQ = QT * G;
  const size_t smem = sizeof(float) * (size_t)kGroups * NQ * (kD + 2);
  dim3 grid(tiles, p.Hkv, p.nsplit);
  MTTS_CUDA_CHECK(mtts_launch(gqa_attn_kernel<QT, G>, grid, dim3(kThreads), smem, stream, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

template <int QT, int G>
int configure_attn() {
  const size_t smem = sizeof(float) * (size_t)kGroups * QT * G * (kD + 2);
  if (smem > 48 * 1024)
    MTTS_CUDA_CHECK(cudaFuncSetAttribute(gqa_attn_kernel<QT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return MTTS_OK;
}

}  // namespace

int mtts_configure_attention() {
  int rc = 0;
  if ((rc = configure_attn<1, 1>())) return rc;
  if ((rc = configure_attn<1, 2>())) return rc;
  if ((rc = configure_attn<1, 4>())) return rc;
  if ((rc = configure_attn<4, 1>())) return rc;
  if ((rc = configure_attn<4, 2>())) return rc;
  if ((rc = configure_attn<4, 4>())) return rc;
  return MTTS_OK;
}

static constexpr size_t kAttnCounterBytes = 65536;

extern "C" size_t mtts_gqa_attention_workspace_bytes(int tiles, int num_kv_heads, int group, int rows_per_tile,
                                                     int nsplit) {
  if (nsplit <= 1) return kAttnCounterBytes;
  return kAttnCounterBytes + (size_t)tiles * num_kv_heads * nsplit * rows_per_tile * group * (kD + 2) * sizeof(float);
}

extern "C" int mtts_gqa_attention(const void* q, const void* k_pool, const void* v_pool, const int* block_table,
                                  int max_pages, int page_size, const int* tile_row0, const int* tile_nrows,
                                  const int* row_seq, const int* positions, void* out, int tiles, int rows_per_tile,
                                  int num_q_heads, int num_kv_heads, int head_dim, int nsplit, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kD, "mtts_gqa_attention: head_dim must be 128 (got %d)", head_dim);
  MTTS_REQUIRE(num_kv_heads > 0 && num_q_heads % num_kv_heads == 0, "mtts_gqa_attention: bad head counts");
  const int G = num_q_heads / num_kv_heads;
  MTTS_REQUIRE(G == 1 || G == 2 || G == 4, "mtts_gqa_attention: GQA group must be 1, 2 or 4 (got %d)", G);
  MTTS_REQUIRE(rows_per_tile == 1 || rows_per_tile == 4, "mtts_gqa_attention: rows_per_tile must be 1 or 4");
  MTTS_REQUIRE(page_size > 0 && (page_size & (page_size - 1)) == 0, "mtts_gqa_attention: page_size must be a power of two");
  MTTS_REQUIRE(nsplit >= 1 && nsplit <= 64, "mtts_gqa_attention: nsplit out of range");
  if (tiles <= 0) return MTTS_OK;
  MTTS_REQUIRE(q && k_pool && v_pool && positions && out, "mtts_gqa_attention: null pointer");
  AttnParams p;
  p.q = reinterpret_cast<const bf16*>(q); p.k_pool = reinterpret_cast<const bf16*>(k_pool);
  p.v_pool = reinterpret_cast<const bf16*>(v_pool); p.block_table = block_table; p.max_pages = max_pages;
  int shift = 0;
  while ((1 << shift) < page_size) ++shift;
  p.page_shift = shift; p.tile_row0 = tile_row0; p.tile_nrows = tile_nrows; p.row_seq = row_seq; p.positions = positions;
  p.out = reinterpret_cast<bf16*>(out); p.Hq = num_q_heads; p.Hkv = num_kv_heads;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  p.nsplit = nsplit; p.ws = nullptr; p.counters = nullptr;
  if (nsplit > 1) {
    const size_t need = mtts_gqa_attention_workspace_bytes(tiles, num_kv_heads, G, rows_per_tile, nsplit);
    MTTS_REQUIRE(workspace && workspace_bytes >= need, "mtts_gqa_attention: workspace too small (%zu < %zu)",
                 workspace_bytes, need);
    MTTS_REQUIRE((size_t)tiles * num_kv_heads * sizeof(int) <= kAttnCounterBytes,
                 "mtts_gqa_attention: too many tiles for split-KV");
    p.counters = reinterpret_cast<int*>(workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kAttnCounterBytes);
  }
  if (rows_per_tile == 1 && tile_row0 == nullptr) {
    if (G == 1) return launch_decode<1>(p, tiles, stream);
    if (G == 2) return launch_decode<2>(p, tiles, stream);
    return launch_decode<4>(p, tiles, stream);
  }
#define MTTS_ATTN_CASE(QT_, G_) \
  if (rows_per_tile == QT_ && G == G_) return launch_attn<QT_, G_>(p, tiles, stream);
  MTTS_ATTN_CASE(1, 1)
  MTTS_ATTN_CASE(1, 2)
  MTTS_ATTN_CASE(1, 4)
  MTTS_ATTN_CASE(4, 1)
  MTTS_ATTN_CASE(4, 2)
  MTTS_ATTN_CASE(4, 4)
#undef MTTS_ATTN_CASE
  return mtts_set_error(MTTS_ERR_UNSUPPORTED, "mtts_gqa_attention: unsupported configuration");
}
