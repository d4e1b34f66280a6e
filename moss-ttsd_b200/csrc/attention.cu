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
#include <stdlib.h>

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
  // fused decode (mtts_gqa_decode_fused): q/k/v of the new row come straight from the projection output
  const bf16* qkv;        // [rows, (Hq + 2 Hkv) * 128]
  long long ld_qkv;
  // ... or as fp32 split-K partial tiles [splits][rows][(Hq + 2 Hkv) * 128] of mtts_gemm_splitk, summed here in
  // ascending slice order and rounded to bf16 once (== mtts_splitk_reduce followed by the bf16 path above)
  const float* qkv_partials;
  int qkv_splits;
  const bf16* q_norm_w;
  const bf16* k_norm_w;
  const float* inv_freq;
  float eps;
  int num_pages;
  int* err_flag;
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

// ------------------------------------------------------------------------------------------------
// Decode on tensor cores. The CUDA-core kernel above needs 2 FMA + an unpack per K/V byte, which is ~13 us of issue
// time per layer at batch 64 (120 MB of K/V) against the 18 us the bytes take to arrive: it is half issue-bound.
// Here the G query heads of a kv head are rows of an m16n8k16 A operand (the other rows are zero), K tiles of 64 keys
// are the B operand of S = Q K^T straight from their row-major shared-memory image (ldmatrix), the probabilities are
// re-packed to bf16 in registers (as the reference's eager path does before P @ V) and V is the B operand of O += P V
// through ldmatrix.trans. K/V tiles arrive by cp.async (16 B per thread, XOR-swizzled 16-byte chunks so that both
// ldmatrix flavours are conflict-free). One CTA = (row, kv head, split), 4 warps x 16 keys per tile; STAGES = 1 when
// enough CTAs share an SM to overlap each other's loads, 2 otherwise.
// ------------------------------------------------------------------------------------------------
constexpr int kTcKeys = 64;
constexpr int kTcTileBytes = kTcKeys * kD * 2;  // 16 KB per K or V tile
constexpr int kTcNewBytes = (8 + 2) * kD * 2;    // fused variant: q heads, k, v of the new row

// 16-byte async copy; `valid == false` writes 16 zero bytes instead (src-size 0), the address is not dereferenced
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// FUSED: the kernel also does what mtts_qknorm_rope_kvappend does for its (row, kv head): per-head RMSNorm + RoPE of the
// G query heads and of the new key straight from the q/k/v projection output, append of the new key/value to the cache
// (by the split that owns position pos) and their insertion into the shared-memory tile, so that neither q nor the new
// K/V row makes a round trip through global memory and the first K/V tile is requested BEFORE the wait on the
// projection kernel (programmatic dependent launch): the cached keys do not depend on it.
template <int G, int STAGES, bool FUSED>
__global__ void __launch_bounds__(128, STAGES == 1 ? 4 : 3) gqa_decode_tc_kernel(const AttnParams p) {
  static_assert(G <= 8, "the q heads of one kv head are rows 0..G-1 of the 16-row MMA tile");
  pdl_launch_dependents();
  if (!FUSED) pdl_wait();
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ int s_is_last;
  const int row = blockIdx.x, hk = blockIdx.y, split = blockIdx.z;
  const int seq = p.row_seq ? p.row_seq[row] : row;
  const int kv_len = p.positions[row] + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  int k_begin = 0, k_end = kv_len;
  if (p.nsplit > 1) {
    const int per = ((kv_len + p.nsplit - 1) / p.nsplit + kTcKeys - 1) / kTcKeys * kTcKeys;
    k_begin = min(kv_len, split * per);
    k_end = min(kv_len, k_begin + per);
  }
  const int ntiles = (k_end - k_begin + kTcKeys - 1) / kTcKeys;

  const int page_mask = (1 << p.page_shift) - 1;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)p.Hkv << p.page_shift) * kD;
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(smraw));

  auto issue = [&](int tile, int stage) {
    const int key0 = k_begin + tile * kTcKeys;
    const uint32_t kdst = smem0 + stage * 2 * kTcTileBytes, vdst = kdst + kTcTileBytes;
    const int c = tid & 15;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kr = (tid >> 4) + 8 * j;                       // key row inside the tile
      // rows past the range are ZERO-filled (their scores are masked, but 0 x stale NaN bits in V would still poison
      // O); in the fused variant so is the row of the new token, which is patched in from shared memory later
      const bool live = key0 + kr < (FUSED ? min(k_end, kv_len - 1) : k_end);
      const int key = min(key0 + kr, k_end - 1);
      const int lp = key >> p.page_shift;
      const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
      const long long off = (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + c * 8;
      const uint32_t so = kr * 256 + ((c ^ (kr & 7)) << 4);    // 16-byte chunks XOR-swizzled by the row
      cp_async16(kdst + so, p.k_pool + off, live);
      cp_async16(vdst + so, p.v_pool + off, live);
    }
    cp_async_commit();
  };

  // O is accumulated TRANSPOSED (O^T = V^T P^T: 16 dims x 8 heads per MMA tile, 8 tiles): 32 accumulator registers
  // instead of the 64 of O = P V, of which half would belong to the zero rows; this is what lets 5 CTAs share an SM.
  // ot[dt][0..1] = dims 16 dt + g, heads 2t, 2t+1 ; ot[dt][2..3] = dims 16 dt + g + 8.
  float ot[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d)
#pragma unroll
    for (int i = 0; i < 4; ++i) ot[d][i] = 0.f;
  float m = -1e30f, l = 0.f;  // online-softmax state of head g (replicated over the 4 lanes of a quad)

  // STAGES == 1 relies on the other CTAs of the SM (4 resident) to keep loads in flight while this one computes.
  if (ntiles > 0) issue(0, 0);
  if (STAGES == 2 && ntiles > 1) issue(1, 1);

  // Q as A fragments: row g = head g (zero for g >= G), 8 k-steps of 16 dims; rows 8..15 of the tile are all zero
  uint32_t qa[8][2];
  bf16* s_new = reinterpret_cast<bf16*>(smraw + STAGES * 2 * kTcTileBytes);  // FUSED: [G q heads | k | v] x 128 bf16
  const int pos = kv_len - 1;
  if (FUSED) {
    pdl_wait();  // the projection output is needed from here on
    const bool owner = pos >= k_begin && pos < k_end;
    for (int item = warp; item < G + 2; item += 4) {  // G query heads, then k, then v of this kv head
      const int col = item < G ? (hk * G + item) * kD : (item == G ? p.Hq * kD + hk * kD : (p.Hq + p.Hkv) * kD + hk * kD);
      float x0, x1, x2, x3;
      if (p.qkv_partials) {
        const long long nq = (long long)(p.Hq + 2 * p.Hkv) * kD;
        const long long slice = (long long)gridDim.x * nq;   // rows x columns of one partial tile (grid.x == rows)
        const float* P = p.qkv_partials + (long long)row * nq + col + 2 * lane;
        float2 a = make_float2(0.f, 0.f), b = a;
#pragma unroll 4
        for (int sidx = 0; sidx < p.qkv_splits; ++sidx) {
          const float2 u = __ldcg(reinterpret_cast<const float2*>(P + sidx * slice));
          const float2 v = __ldcg(reinterpret_cast<const float2*>(P + sidx * slice + 64));
          a.x += u.x; a.y += u.y; b.x += v.x; b.y += v.y;
        }
        x0 = bf16_round(a.x); x1 = bf16_round(a.y); x2 = bf16_round(b.x); x3 = bf16_round(b.y);
      } else {
        const bf16* src = p.qkv + (long long)row * p.ld_qkv + col;
        const uint32_t a = *reinterpret_cast<const uint32_t*>(src + 2 * lane);       // elements 2l, 2l+1
        const uint32_t b = *reinterpret_cast<const uint32_t*>(src + 64 + 2 * lane);  // elements 64+2l, 65+2l
        x0 = bf16lo(a); x1 = bf16hi(a); x2 = bf16lo(b); x3 = bf16hi(b);
      }
      if (item <= G) {  // same arithmetic, in the same order, as qknorm_rope_kv_kernel (lm_ops.cu)
        const bf16* nw = item < G ? p.q_norm_w : p.k_norm_w;
        float ss = x0 * x0;
        ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
        ss = warp_sum(ss);
        const float inv = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
        const uint32_t wa = *reinterpret_cast<const uint32_t*>(nw + 2 * lane);
        const uint32_t wb = *reinterpret_cast<const uint32_t*>(nw + 64 + 2 * lane);
        x0 = bf16_round(bf16lo(wa) * bf16_round(x0 * inv));
        x1 = bf16_round(bf16hi(wa) * bf16_round(x1 * inv));
        x2 = bf16_round(bf16lo(wb) * bf16_round(x2 * inv));
        x3 = bf16_round(bf16hi(wb) * bf16_round(x3 * inv));
        const float f0 = (float)pos * p.inv_freq[2 * lane];
        const float f1 = (float)pos * p.inv_freq[2 * lane + 1];
        const float c0 = bf16_round(cosf(f0)), s0 = bf16_round(sinf(f0));
        const float c1 = bf16_round(cosf(f1)), s1 = bf16_round(sinf(f1));
        const float o0 = bf16_round(x0 * c0) + bf16_round(-x2 * s0);
        const float o1 = bf16_round(x1 * c1) + bf16_round(-x3 * s1);
        const float o2 = bf16_round(x2 * c0) + bf16_round(x0 * s0);
        const float o3 = bf16_round(x3 * c1) + bf16_round(x1 * s1);
        x0 = o0; x1 = o1; x2 = o2; x3 = o3;
      }
      const uint32_t lo = pack_bf16(x0, x1), hi = pack_bf16(x2, x3);
      *reinterpret_cast<uint32_t*>(s_new + item * kD + 2 * lane) = lo;
      *reinterpret_cast<uint32_t*>(s_new + item * kD + 64 + 2 * lane) = hi;
      if (item >= G && owner) {  // append the new key / value row to the cache
        const int lp = pos >> p.page_shift;
        int page = -1;
        if (pos >= 0 && lp < p.max_pages) page = p.block_table ? p.block_table[(long long)seq * p.max_pages + lp] : seq * p.max_pages + lp;
        if (page < 0 || page >= p.num_pages) {
          if (p.err_flag && lane == 0) *p.err_flag = 2;
        } else {
          bf16* dst = const_cast<bf16*>(item == G ? p.k_pool : p.v_pool) + (long long)page * page_stride + head_off +
                      (long long)(pos & page_mask) * kD;
          *reinterpret_cast<uint32_t*>(dst + 2 * lane) = lo;
          *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = hi;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    qa[ks][0] = qa[ks][1] = 0u;
    if (g < G) {
      const bf16* qp = FUSED ? s_new + g * kD + ks * 16 + 2 * t : p.q + ((long long)row * p.Hq + hk * G + g) * kD + ks * 16 + 2 * t;
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qp);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qp + 8);
    }
  }
#pragma unroll 1
  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = STAGES == 2 ? (tile & 1) : 0;
    if (STAGES == 2 && tile + 1 < ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const uint32_t kbase = smem0 + stage * 2 * kTcTileBytes, vbase = kbase + kTcTileBytes;
    if (FUSED) {  // the new row is not in the cache image that was copied: put it into the tile (uniform branch)
      const int r = pos - (k_begin + tile * kTcKeys);
      if (r >= 0 && r < kTcKeys) {
        if (tid < 32) {
          const int c = tid & 15;
          const uint4 v = *reinterpret_cast<const uint4*>(s_new + (G + (tid >> 4)) * kD + c * 8);
          uint8_t* dst = smraw + stage * 2 * kTcTileBytes + (tid >> 4) * kTcTileBytes + r * 256 + ((c ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = v;
        }
        __syncthreads();
      }
    }
    // ---- S = Q K^T for this warp's 16 keys (two 8-key column tiles)
    float s[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
    {
      const int kr = 16 * warp + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int ch = 2 * ks + ((lane >> 3) & 1);
        uint32_t b[4];
        ldsm_x4(b, kbase + kr * 256 + ((ch ^ (kr & 7)) << 4));
        mma_bf16(s[0], qa[ks][0], 0u, qa[ks][1], 0u, b[0], b[1]);
        mma_bf16(s[1], qa[ks][0], 0u, qa[ks][1], 0u, b[2], b[3]);
      }
    }
    // ---- online softmax of row g over these 16 keys (this lane holds keys 2t, 2t+1 of both column tiles)
    const int kcol = k_begin + tile * kTcKeys + 16 * warp + 2 * t;
    float sc[4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) sc[2 * j + e] = (kcol + 8 * j + e < k_end) ? s[j][e] * p.scale_log2 : -INFINITY;
    float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float mn = fmaxf(m, mx);
    const float corr = fast_exp2(m - mn);
    m = mn;
    float ps = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc[i] = fast_exp2(sc[i] - mn);
      ps += sc[i];
    }
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    l = fmaf(l, corr, ps);
    {
      // this lane's accumulator columns are heads 2t and 2t+1: their rescale factors live in lanes 8t and 8t+4
      const float c0 = __shfl_sync(0xffffffffu, corr, 8 * t), c1 = __shfl_sync(0xffffffffu, corr, 8 * t + 4);
#pragma unroll
      for (int d = 0; d < 8; ++d) { ot[d][0] *= c0; ot[d][1] *= c1; ot[d][2] *= c0; ot[d][3] *= c1; }
    }
    // ---- O^T += V^T P^T : V^T tiles (16 dims x 16 keys) are the A operand through ldmatrix.trans; P^T (bf16, as the
    // reference casts the probabilities) is the B operand, and the score accumulators already have its layout
    // (k = keys 2t, 2t+1 (+8), n = head g)
    const uint32_t pb0 = pack_bf16(sc[0], sc[1]), pb1 = pack_bf16(sc[2], sc[3]);
    {
      const int kr = 16 * warp + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const int ch = 2 * dt + ((lane >> 3) & 1);
        uint32_t a[4];
        ldsm_x4_trans(a, vbase + kr * 256 + ((ch ^ (kr & 7)) << 4));
        mma_bf16(ot[dt], a[0], a[1], a[2], a[3], pb0, pb1);
      }
    }
    if (tile + STAGES < ntiles) {
      __syncthreads();  // every warp is done with this stage
      issue(tile + STAGES, stage);
    }
  }

  // ---- combine the 4 warps (rows g < G live in lanes 0 .. 4G-1), then the splits
  __syncthreads();
  float* sm_m = reinterpret_cast<float*>(smraw);      // [4][G]
  float* sm_l = sm_m + 4 * G;                          // [4][G]
  float* sm_o = sm_l + 4 * G;                          // [4][G][128]
  if (g < G && t == 0) { sm_m[warp * G + g] = m; sm_l[warp * G + g] = l; }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int head = 2 * t + e;
    if (head < G) {
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        sm_o[(warp * G + head) * kD + 16 * d + g] = ot[d][e];
        sm_o[(warp * G + head) * kD + 16 * d + g + 8] = ot[d][2 + e];
      }
    }
  }
  __syncthreads();
  const int unit = row * p.Hkv + hk;
  float* wsb = p.ws ? p.ws + ((long long)unit * p.nsplit) * G * (kD + 2) : nullptr;
  for (int oi = tid; oi < G * kD; oi += 128) {
    const int gg = oi / kD, d = oi % kD;
    float M = -1e30f;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, sm_m[w * G + gg]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float wt = fast_exp2(sm_m[w * G + gg] - M);
      L = fmaf(sm_l[w * G + gg], wt, L);
      O = fmaf(sm_o[(w * G + gg) * kD + d], wt, O);
    }
    if (p.nsplit == 1) {
      p.out[((long long)row * p.Hq + hk * G + gg) * kD + d] = __float2bfloat16_rn(O / L);
    } else {
      float* w0 = wsb + (long long)split * G * (kD + 2) + gg * (kD + 2);
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
  if (tid == 0) {
    const int prev = atomicAdd(p.counters + unit, 1);
    const int last = prev == p.nsplit - 1;
    if (last) p.counters[unit] = 0;
    s_is_last = last;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  for (int oi = tid; oi < G * kD; oi += 128) {
    const int gg = oi / kD, d = oi % kD;
    float M = -1e30f;
    for (int sidx = 0; sidx < p.nsplit; ++sidx) M = fmaxf(M, __ldcg(wsb + ((long long)sidx * G + gg) * (kD + 2)));
    float L = 0.f, O = 0.f;
    for (int sidx = 0; sidx < p.nsplit; ++sidx) {
      const float* w0 = wsb + ((long long)sidx * G + gg) * (kD + 2);
      const float wt = fast_exp2(__ldcg(w0) - M);
      L = fmaf(__ldcg(w0 + 1), wt, L);
      O = fmaf(__ldcg(w0 + 2 + d), wt, O);
    }
    p.out[((long long)row * p.Hq + hk * G + gg) * kD + d] = __float2bfloat16_rn(O / L);
  }
}

template <int G, bool FUSED>
int launch_decode_tc(const AttnParams& p, int rows, cudaStream_t stream) {
  dim3 grid(rows, p.Hkv, p.nsplit);
  const long long ctas = (long long)rows * p.Hkv * p.nsplit;
  static int force_stages = -1;
  if (force_stages < 0) {
    const char* e = getenv("MTTS_ATTN_STAGES");
    force_stages = e ? atoi(e) : 0;
  }
  // One K/V stage: 4 CTAs per SM; two stages: 3 per SM but loads overlap inside the CTA. Time goes with the number of
  // CTA waves, so the variant that fills its last wave better wins (batch 64: 512 CTAs on 592 slots -> one stage,
  // 25.6 vs 31.4 us; batch 96: 768 CTAs -> two stages, 45.8 vs 51.9 us); comparable fills favour two stages.
  const long long slots1 = 4LL * mtts_num_sms(), slots2 = 3LL * mtts_num_sms();
  const double fill1 = (double)ctas / (double)(((ctas + slots1 - 1) / slots1) * slots1);
  const double fill2 = (double)ctas / (double)(((ctas + slots2 - 1) / slots2) * slots2);
  const bool one_stage = force_stages ? force_stages == 1 : fill2 < 0.97 * fill1;
  if (one_stage) {
    MTTS_CUDA_CHECK(mtts_launch(gqa_decode_tc_kernel<G, 1, FUSED>, grid, dim3(128), (size_t)2 * kTcTileBytes + kTcNewBytes, stream, p));
  } else {
    MTTS_CUDA_CHECK(mtts_launch(gqa_decode_tc_kernel<G, 2, FUSED>, grid, dim3(128), (size_t)4 * kTcTileBytes + kTcNewBytes, stream, p));
  }
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// Prefill on tensor cores: bf16 mma.sync m16n8k16 with fp32 accumulation, flash-style online softmax.
// CTA = up to 64 consecutive query rows of ONE sequence x one q head (4 warps x 16 rows); K tiles of 64 keys are
// staged row-major ([key][d], pitch 136) and V tiles TRANSPOSED ([d][key], pitch 72) so that every B fragment is one
// conflict-free 32-bit shared load. P keeps the accumulator layout, which for k16 is exactly the A-fragment layout
// of two adjacent 8-key blocks. Causal: row (pos0 + r) sees keys <= its own position.
// ------------------------------------------------------------------------------------------------
constexpr int kPQ = 64, kPK = 64, kKPitch = 136, kVPitch = 72;

__device__ __forceinline__ void mma_bf16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------
// Prefill, flash style on the same machinery as the decode kernel above: CTA = up to 64 consecutive query rows of one
// sequence x one q head (4 warps x 16 rows, every MMA row is a real query), K/V tiles of 64 keys double-buffered by
// cp.async with XOR-swizzled 16-byte chunks, K as the B operand of S = Q K^T through ldmatrix, V as the B operand of
// O += P V through ldmatrix.trans, P re-packed to bf16 in registers. Causal: row (pos0 + r) sees keys <= pos0 + r.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3) gqa_prefill_fa_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tile = blockIdx.x, hq = blockIdx.y;
  const int G = p.Hq / p.Hkv, hk = hq / G;
  const int row0 = p.tile_row0[tile], nrows = p.tile_nrows[tile];
  const int seq = p.row_seq ? p.row_seq[row0] : row0;
  const int pos0 = p.positions[row0];
  const int kv_len = pos0 + nrows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (kv_len + kTcKeys - 1) / kTcKeys;
  const int page_mask = (1 << p.page_shift) - 1;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)p.Hkv << p.page_shift) * kD;
  const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(smraw));

  auto issue = [&](int kt, int stage) {
    const int key0 = kt * kTcKeys;
    const uint32_t kdst = smem0 + stage * 2 * kTcTileBytes, vdst = kdst + kTcTileBytes;
    const int c = tid & 15;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kr = (tid >> 4) + 8 * j;
      const bool live = key0 + kr < kv_len;  // rows past the sequence are zero-filled
      const int key = min(key0 + kr, kv_len - 1);
      const int lp = key >> p.page_shift;
      const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
      const long long off = (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + c * 8;
      const uint32_t so = kr * 256 + ((c ^ (kr & 7)) << 4);
      cp_async16(kdst + so, p.k_pool + off, live);
      cp_async16(vdst + so, p.v_pool + off, live);
    }
    cp_async_commit();
  };
  issue(0, 0);
  if (ntiles > 1) issue(1, 1);

  // Q fragments: rows r0 = 16 warp + g and r0 + 8 (clamped for the load, masked at the store), 8 k-steps of 16 dims
  uint32_t qa[8][4];
  {
    const int ra = min(warp * 16 + g, nrows - 1), rb = min(warp * 16 + g + 8, nrows - 1);
    const bf16* qpa = p.q + ((long long)(row0 + ra) * p.Hq + hq) * kD + 2 * t;
    const bf16* qpb = p.q + ((long long)(row0 + rb) * p.Hq + hq) * kD + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qpa + ks * 16);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qpb + ks * 16);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(qpa + ks * 16 + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(qpb + ks * 16 + 8);
    }
  }
  float o[16][4];
#pragma unroll
  for (int d = 0; d < 16; ++d)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[d][i] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
  const int qp0 = pos0 + warp * 16 + g, qp1 = qp0 + 8;  // positions of this lane's two query rows

#pragma unroll 1
  for (int kt = 0; kt < ntiles; ++kt) {
    const int stage = kt & 1;
    if (kt + 1 < ntiles) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const uint32_t kbase = smem0 + stage * 2 * kTcTileBytes, vbase = kbase + kTcTileBytes;
    const int key0 = kt * kTcKeys;
    if (key0 <= pos0 + warp * 16 + 15) {  // warp-uniform: some key of this tile is visible to some row of this warp
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          const int kr = 16 * jp + (lane & 7) + ((lane >> 4) & 1) * 8;
          const int ch = 2 * ks + ((lane >> 3) & 1);
          uint32_t b[4];
          ldsm_x4(b, kbase + kr * 256 + ((ch ^ (kr & 7)) << 4));
          mma_bf16(s[2 * jp], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[0], b[1]);
          mma_bf16(s[2 * jp + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[2], b[3]);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = key0 + 8 * j + 2 * t + e;
          s[j][e] = (key <= qp0 && key < kv_len) ? s[j][e] * p.scale_log2 : -INFINITY;
          s[j][2 + e] = (key <= qp1 && key < kv_len) ? s[j][2 + e] * p.scale_log2 : -INFINITY;
          mx0 = fmaxf(mx0, s[j][e]);
          mx1 = fmaxf(mx1, s[j][2 + e]);
        }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s[j][e] = fast_exp2(s[j][e] - mn0);
          s[j][2 + e] = fast_exp2(s[j][2 + e] - mn1);
          ps0 += s[j][e];
          ps1 += s[j][2 + e];
        }
      ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1);
      ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
      ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1);
      ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
      l0 = fmaf(l0, c0, ps0);
      l1 = fmaf(l1, c1, ps1);
#pragma unroll
      for (int d = 0; d < 16; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 16 keys per k-step: score column tiles 2kk, 2kk+1 are the A fragment
        const uint32_t pa0 = pack_bf16(s[2 * kk][0], s[2 * kk][1]), pa1 = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        const uint32_t pa2 = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pa3 = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int kr = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
          const int ch = 2 * dp + ((lane >> 4) & 1);
          uint32_t b[4];
          ldsm_x4_trans(b, vbase + kr * 256 + ((ch ^ (kr & 7)) << 4));
          mma_bf16(o[2 * dp], pa0, pa1, pa2, pa3, b[0], b[1]);
          mma_bf16(o[2 * dp + 1], pa0, pa1, pa2, pa3, b[2], b[3]);
        }
      }
    }
    if (kt + 2 < ntiles) {
      __syncthreads();  // every warp is done with this stage
      issue(kt + 2, stage);
    }
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int ra = warp * 16 + g, rb = ra + 8;
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    if (ra < nrows)
      *reinterpret_cast<uint32_t*>(p.out + ((long long)(row0 + ra) * p.Hq + hq) * kD + 8 * d + 2 * t) = pack_bf16(o[d][0] * i0, o[d][1] * i0);
    if (rb < nrows)
      *reinterpret_cast<uint32_t*>(p.out + ((long long)(row0 + rb) * p.Hq + hq) * kD + 8 * d + 2 * t) = pack_bf16(o[d][2] * i1, o[d][3] * i1);
  }
}

__global__ void __launch_bounds__(128) gqa_prefill_tc_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sk = reinterpret_cast<bf16*>(smraw);        // [64 keys][136]
  bf16* svt = sk + kPK * kKPitch;                   // [128 d][72]
  const int tile = blockIdx.x, hq = blockIdx.y;
  const int G = p.Hq / p.Hkv;
  const int hk = hq / G;
  const int row0 = p.tile_row0[tile];
  const int nrows = p.tile_nrows[tile];
  const int seq = p.row_seq ? p.row_seq[row0] : row0;
  const int pos0 = p.positions[row0];
  const int kv_len = pos0 + nrows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = p.scale_log2;

  // Q fragments: rows (warp*16 + g) and (+8), 8 k-blocks of 16 dims
  uint32_t qa[8][4];
  {
    const int r0 = min(warp * 16 + g, nrows - 1), r1 = min(warp * 16 + g + 8, nrows - 1);
    const bf16* q0p = p.q + ((long long)(row0 + r0) * p.Hq + hq) * kD;
    const bf16* q1p = p.q + ((long long)(row0 + r1) * p.Hq + hq) * kD;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16 + 2 * t);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16 + 2 * t);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16 + 2 * t + 8);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16 + 2 * t + 8);
    }
  }
  float o[16][4];
#pragma unroll
  for (int nb = 0; nb < 16; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nb][i] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
  const int qpos0 = pos0 + warp * 16 + g, qpos1 = qpos0 + 8;  // positions of this thread's two rows

  const int page_mask = (1 << p.page_shift) - 1;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)p.Hkv << p.page_shift) * kD;

  for (int k0 = 0; k0 < kv_len; k0 += kPK) {
    __syncthreads();
    // stage K [key][d] and V^T [d][key]: 64 keys x 16 chunks of 8 dims
    for (int i = tid; i < kPK * 16; i += 128) {
      const int r = i >> 4, c8 = (i & 15) * 8;
      const int key = k0 + r;
      uint4 kq = make_uint4(0, 0, 0, 0), vq = kq;
      if (key < kv_len) {
        const int lp = key >> p.page_shift;
        const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
        const long long off = (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + c8;
        kq = *reinterpret_cast<const uint4*>(p.k_pool + off);
        vq = *reinterpret_cast<const uint4*>(p.v_pool + off);
      }
      *reinterpret_cast<uint4*>(sk + r * kKPitch + c8) = kq;
      const bf16* ve = reinterpret_cast<const bf16*>(&vq);
#pragma unroll
      for (int e = 0; e < 8; ++e) svt[(c8 + e) * kVPitch + r] = ve[e];
    }
    __syncthreads();
    if (k0 > pos0 + warp * 16 + 15) continue;  // this warp's rows see none of these keys (warp-uniform)
    // ---- S = Q K^T
    float sc[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[nb][i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sk + (nb * 8 + g) * kKPitch + kk * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sk + (nb * 8 + g) * kKPitch + kk * 16 + 2 * t + 8);
        mma_bf16_16x8x16(sc[nb], qa[kk], b0, b1);
      }
    }
    // ---- scale, causal mask, online softmax
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int key = k0 + nb * 8 + 2 * t;
      sc[nb][0] = (key <= qpos0) ? sc[nb][0] * scale_log2 : -INFINITY;
      sc[nb][1] = (key + 1 <= qpos0) ? sc[nb][1] * scale_log2 : -INFINITY;
      sc[nb][2] = (key <= qpos1) ? sc[nb][2] * scale_log2 : -INFINITY;
      sc[nb][3] = (key + 1 <= qpos1) ? sc[nb][3] * scale_log2 : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(sc[nb][0], sc[nb][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nb][2], sc[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = fast_exp2(m0 - mx0), c1 = fast_exp2(m1 - mx1);
    m0 = mx0;
    m1 = mx1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      sc[nb][0] = fast_exp2(sc[nb][0] - mx0);
      sc[nb][1] = fast_exp2(sc[nb][1] - mx0);
      sc[nb][2] = fast_exp2(sc[nb][2] - mx1);
      sc[nb][3] = fast_exp2(sc[nb][3] - mx1);
      ps0 += sc[nb][0] + sc[nb][1];
      ps1 += sc[nb][2] + sc[nb][3];
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
      o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
    }
    // ---- O += P V : 4 key blocks of 16, 16 d blocks of 8
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t pa[4];
      pa[0] = pack_bf16(sc[2 * j][0], sc[2 * j][1]);
      pa[1] = pack_bf16(sc[2 * j][2], sc[2 * j][3]);
      pa[2] = pack_bf16(sc[2 * j + 1][0], sc[2 * j + 1][1]);
      pa[3] = pack_bf16(sc[2 * j + 1][2], sc[2 * j + 1][3]);
#pragma unroll
      for (int nb = 0; nb < 16; ++nb) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(svt + (nb * 8 + g) * kVPitch + j * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(svt + (nb * 8 + g) * kVPitch + j * 16 + 2 * t + 8);
        mma_bf16_16x8x16(o[nb], pa, b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    if (r0 < nrows)
      *reinterpret_cast<uint32_t*>(p.out + ((long long)(row0 + r0) * p.Hq + hq) * kD + nb * 8 + 2 * t) =
          pack_bf16(o[nb][0] * i0, o[nb][1] * i0);
    if (r1 < nrows)
      *reinterpret_cast<uint32_t*>(p.out + ((long long)(row0 + r1) * p.Hq + hq) * kD + nb * 8 + 2 * t) =
          pack_bf16(o[nb][2] * i1, o[nb][3] * i1);
  }
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
  const int big = 4 * kTcTileBytes + kTcNewBytes;
  if (cudaFuncSetAttribute(gqa_prefill_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kTcTileBytes) != cudaSuccess)
    return mtts_set_error(MTTS_ERR_CUDA, "attention: smem attribute (prefill)");
  cudaError_t e = cudaFuncSetAttribute(gqa_decode_tc_kernel<1, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gqa_decode_tc_kernel<2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gqa_decode_tc_kernel<4, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gqa_decode_tc_kernel<1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gqa_decode_tc_kernel<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gqa_decode_tc_kernel<4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e != cudaSuccess) return mtts_set_error(MTTS_ERR_CUDA, "attention: smem attribute: %s", cudaGetErrorString(e));
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
  MTTS_REQUIRE(rows_per_tile == 1 || rows_per_tile == 4 || rows_per_tile == 64,
               "mtts_gqa_attention: rows_per_tile must be 1, 4 or 64");
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
  if (rows_per_tile == 64) {
    MTTS_REQUIRE(tile_row0 && tile_nrows, "mtts_gqa_attention: 64-row tiles need tile_row0 / tile_nrows");
    MTTS_REQUIRE(nsplit == 1, "mtts_gqa_attention: prefill tiles do not split the keys");
    static int use_fa = -1;
    if (use_fa < 0) {
      const char* e = getenv("MTTS_PREFILL_OLD");
      use_fa = (e && e[0] == '1') ? 0 : 1;
    }
    if (use_fa) {
      MTTS_CUDA_CHECK(mtts_launch(gqa_prefill_fa_kernel, dim3(tiles, num_q_heads), dim3(128), (size_t)4 * kTcTileBytes, stream, p));
      MTTS_LAUNCH_CHECK();
      return MTTS_OK;
    }
    const size_t smem = sizeof(bf16) * (kPK * kKPitch + kD * kVPitch);
    MTTS_CUDA_CHECK(mtts_launch(gqa_prefill_tc_kernel, dim3(tiles, num_q_heads), dim3(128), smem, stream, p));
    MTTS_LAUNCH_CHECK();
    return MTTS_OK;
  }
  if (rows_per_tile == 1 && tile_row0 == nullptr) {
    static int use_tc = -1;
    if (use_tc < 0) {
      const char* e = getenv("MTTS_ATTN_SIMT");
      use_tc = (e && e[0] == '1') ? 0 : 1;
    }
    if (use_tc) {
      if (G == 1) return launch_decode_tc<1, false>(p, tiles, stream);
      if (G == 2) return launch_decode_tc<2, false>(p, tiles, stream);
      if (G == 4) return launch_decode_tc<4, false>(p, tiles, stream);
    }
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

static int gqa_decode_fused_impl(const void* qkv, long long ld_qkv, const float* qkv_partials, int qkv_splits, const void* q_norm_w, const void* k_norm_w,
                                     const float* inv_freq, float eps, void* k_pool, void* v_pool, const int* block_table,
                                     int max_pages, int page_size, int num_pages, const int* positions, void* out, int rows,
                                     int num_q_heads, int num_kv_heads, int head_dim, int nsplit, void* workspace,
                                     size_t workspace_bytes, int* err_flag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kD, "mtts_gqa_decode_fused: head_dim must be 128 (got %d)", head_dim);
  MTTS_REQUIRE(page_size > 0 && (page_size & (page_size - 1)) == 0, "mtts_gqa_decode_fused: page_size must be a power of two");
  MTTS_REQUIRE(num_kv_heads > 0 && num_q_heads % num_kv_heads == 0, "mtts_gqa_decode_fused: bad head counts");
  const int G = num_q_heads / num_kv_heads;
  MTTS_REQUIRE(G == 1 || G == 2 || G == 4, "mtts_gqa_decode_fused: q heads per kv head must be 1, 2 or 4 (got %d)", G);
  MTTS_REQUIRE(nsplit >= 1 && nsplit <= 64, "mtts_gqa_decode_fused: nsplit out of range");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE((qkv || qkv_partials) && q_norm_w && k_norm_w && inv_freq && k_pool && v_pool && positions && out,
               "mtts_gqa_decode_fused: null pointer");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = nullptr;
  p.k_pool = reinterpret_cast<const bf16*>(k_pool); p.v_pool = reinterpret_cast<const bf16*>(v_pool);
  p.block_table = block_table; p.max_pages = max_pages;
  int shift = 0;
  while ((1 << shift) < page_size) ++shift;
  p.page_shift = shift;
  p.positions = positions; p.out = reinterpret_cast<bf16*>(out);
  p.Hq = num_q_heads; p.Hkv = num_kv_heads;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  p.nsplit = nsplit;
  if (nsplit > 1) {
    const size_t need = mtts_gqa_attention_workspace_bytes(rows, num_kv_heads, G, 1, nsplit);
    MTTS_REQUIRE(workspace && workspace_bytes >= need, "mtts_gqa_decode_fused: workspace too small (%zu < %zu)", workspace_bytes, need);
    p.counters = reinterpret_cast<int*>(workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kAttnCounterBytes);
  }
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.ld_qkv = ld_qkv;
  p.qkv_partials = qkv_partials; p.qkv_splits = qkv_splits;
  p.q_norm_w = reinterpret_cast<const bf16*>(q_norm_w); p.k_norm_w = reinterpret_cast<const bf16*>(k_norm_w);
  p.inv_freq = inv_freq; p.eps = eps; p.num_pages = num_pages; p.err_flag = err_flag;
  if (G == 1) return launch_decode_tc<1, true>(p, rows, stream);
  if (G == 2) return launch_decode_tc<2, true>(p, rows, stream);
  return launch_decode_tc<4, true>(p, rows, stream);
}

extern "C" int mtts_gqa_decode_fused(const void* qkv, long long ld_qkv, const void* q_norm_w, const void* k_norm_w,
                                     const float* inv_freq, float eps, void* k_pool, void* v_pool, const int* block_table,
                                     int max_pages, int page_size, int num_pages, const int* positions, void* out, int rows,
                                     int num_q_heads, int num_kv_heads, int head_dim, int nsplit, void* workspace,
                                     size_t workspace_bytes, int* err_flag, void* stream_) {
  return gqa_decode_fused_impl(qkv, ld_qkv, nullptr, 0, q_norm_w, k_norm_w, inv_freq, eps, k_pool, v_pool, block_table,
                               max_pages, page_size, num_pages, positions, out, rows, num_q_heads, num_kv_heads, head_dim,
                               nsplit, workspace, workspace_bytes, err_flag, stream_);
}

extern "C" int mtts_gqa_decode_fused_splitk(const float* qkv_partials, int qkv_splits, const void* q_norm_w,
                                            const void* k_norm_w, const float* inv_freq, float eps, void* k_pool, void* v_pool,
                                            const int* block_table, int max_pages, int page_size, int num_pages,
                                            const int* positions, void* out, int rows, int num_q_heads, int num_kv_heads,
                                            int head_dim, int nsplit, void* workspace, size_t workspace_bytes, int* err_flag,
                                            void* stream_) {
  MTTS_REQUIRE(qkv_partials != nullptr && qkv_splits >= 1, "mtts_gqa_decode_fused_splitk: needs partial tiles and splits >= 1");
  return gqa_decode_fused_impl(nullptr, 0, qkv_partials, qkv_splits, q_norm_w, k_norm_w, inv_freq, eps, k_pool, v_pool,
                               block_table, max_pages, page_size, num_pages, positions, out, rows, num_q_heads, num_kv_heads,
                               head_dim, nsplit, workspace, workspace_bytes, err_flag, stream_);
}
