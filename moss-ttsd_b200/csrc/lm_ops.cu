// Small bandwidth-bound ops of the Qwen3-style decoder step (bf16 activations, fp32 statistics).
// Every rounding point mirrors the reference's bf16 eager path so that
// teacher-forced logits stay within bf16 noise of it.
#include "common.cuh"
#include "mtts_internal.h"

namespace {

// ------------------------------------------------------------------------------------------------
// 8-table embedding gather + sum.   modeling_asteroid.py:235-250 (_prepare_multi_modal_inputs):
//   acc = zeros(bf16); for i in range(8): acc += E_i[ids[..., i]]      (bf16 rounding after every add)
// ------------------------------------------------------------------------------------------------
struct EmbedParams {
  const long long* ids;  // [rows, channels]
  const bf16* tables[8];
  int vocab[8];
  int rows, channels, hidden;
  bf16* out;  // [rows, hidden]
  int* err_flag;
};

__global__ void embed_sum_kernel(const EmbedParams p) {
  pdl_launch_dependents();  // the GEMM that follows may start streaming its weights now
  pdl_wait();               // launched with the PDL attribute: our own inputs come from the predecessor
  const int vec_per_row = p.hidden >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = gid / vec_per_row;
  if (row >= p.rows) return;
  const int col = static_cast<int>(gid % vec_per_row) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int c = 0; c < p.channels; ++c) {
    const long long id = p.ids[row * p.channels + c];
    if (id < 0 || id >= p.vocab[c]) {
      if (p.err_flag) *p.err_flag = 1;
      continue;
    }
    const uint4 v = *reinterpret_cast<const uint4*>(p.tables[c] + id * p.hidden + col);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] = bf16_round(acc[2 * j] + bf16lo(u[j]));
      acc[2 * j + 1] = bf16_round(acc[2 * j + 1] + bf16hi(u[j]));
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]);
  o.y = pack_bf16(acc[2], acc[3]);
  o.z = pack_bf16(acc[4], acc[5]);
  o.w = pack_bf16(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(p.out + row * p.hidden + col) = o;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm (HF Qwen3RMSNorm, installed modeling_qwen3.py:50-66; used at every decoder layer and as
// the final norm, invoked via modeling_asteroid.py:273-284):
//   v = mean(x.float()^2); y = (x.float() * rsqrt(v + eps)).to(bf16); out = w * y   (bf16 multiply)
// One CTA per row; hidden <= 8 * 1024 elements.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, long long ldx,
                                                      const bf16* __restrict__ w, bf16* __restrict__ out,
                                                      long long ldo, int hidden, float eps) {
  pdl_launch_dependents();  // the GEMM that follows may start streaming its weights now
  pdl_wait();
  __shared__ float red[33];
  const long long row = blockIdx.x;
  const bf16* xr = x + row * ldx;
  const int nvec = hidden >> 3;
  float ss = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + v * 8);
    const uint32_t a[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16lo(a[j]), hi = bf16hi(a[j]);
      ss = fmaf(lo, lo, ss);
      ss = fmaf(hi, hi, ss);
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / (float)hidden + eps);
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + v * 8);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + v * 8);
    const uint32_t a[4] = {u.x, u.y, u.z, u.w};
    const uint32_t b[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16_round(bf16lo(a[j]) * inv) * bf16lo(b[j]);
      const float hi = bf16_round(bf16hi(a[j]) * inv) * bf16hi(b[j]);
      o[j] = pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(out + row * ldo + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Decode-sized variant (hidden == 2048): one WARP per row, the row stays in registers (8 x 16 B per lane), the
// statistic is a shuffle reduction — no shared memory, no block barrier, one pass over the data. 3.9 -> ~2 us per
// launch, and a decode step has 57 of them.
__global__ void __launch_bounds__(256) rmsnorm_warp_kernel(const bf16* __restrict__ x, long long ldx,
                                                           const bf16* __restrict__ w, bf16* __restrict__ out,
                                                           long long ldo, int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16* xr = x + row * ldx;
  uint4 u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t a[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16lo(a[j]), hi = bf16hi(a[j]);
      ss = fmaf(lo, lo, ss);
      ss = fmaf(hi, hi, ss);
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss * (1.0f / 2048.0f) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 wv = *reinterpret_cast<const uint4*>(w + (i * 32 + lane) * 8);
    const uint32_t a[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
    const uint32_t b[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16_round(bf16lo(a[j]) * inv) * bf16lo(b[j]);
      const float hi = bf16_round(bf16hi(a[j]) * inv) * bf16hi(b[j]);
      o[j] = pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(out + row * ldo + (i * 32 + lane) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Consumers of mtts_gemm_splitk's fp32 partial tiles [S][M][N] (gemm_tc.cu): the S slices are summed in ascending
// order (bitwise deterministic) by the kernel that reads the projection anyway.
//   splitk_reduce:          out = bf16(sum_s P_s)                                    (q/k/v in front of attention)
//   splitk_reduce_rmsnorm:  x  <- bf16(x + bf16(sum_s P_s))     residual stream, the reference's two roundings
//                           xn <- RMSNorm(x) * w                (o_proj -> post-attention norm, down_proj -> next input norm)
// One CTA per row; a thread owns 8 consecutive columns per pass.
// ------------------------------------------------------------------------------------------------
// All slice loads of a thread are issued before the first add (the partial tiles sit in L2: ONE round trip, the warp
// issues in order and would otherwise wait for each group before requesting the next); the adds run in ascending slice
// order. S <= 16 (mtts_gemm_splitk's limit): up to 32 float4 in flight per thread.
template <int kMaxS>
__device__ __forceinline__ void sum_slices8_t(const float* __restrict__ P, int S, long long slice, float (&a)[8]) {
  float4 v[kMaxS][2];
#pragma unroll
  for (int i = 0; i < kMaxS; ++i) {
    if (i < S) {
      v[i][0] = __ldcg(reinterpret_cast<const float4*>(P + i * slice));
      v[i][1] = __ldcg(reinterpret_cast<const float4*>(P + i * slice) + 1);
    }
  }
  a[0] = v[0][0].x; a[1] = v[0][0].y; a[2] = v[0][0].z; a[3] = v[0][0].w;
  a[4] = v[0][1].x; a[5] = v[0][1].y; a[6] = v[0][1].z; a[7] = v[0][1].w;
#pragma unroll
  for (int i = 1; i < kMaxS; ++i) {
    if (i < S) {
      a[0] += v[i][0].x; a[1] += v[i][0].y; a[2] += v[i][0].z; a[3] += v[i][0].w;
      a[4] += v[i][1].x; a[5] += v[i][1].y; a[6] += v[i][1].z; a[7] += v[i][1].w;
    }
  }
}
__device__ __forceinline__ void sum_slices8(const float* __restrict__ P, int S, long long slice, float (&a)[8]) {
  if (S <= 4) sum_slices8_t<4>(P, S, slice, a);
  else if (S <= 9) sum_slices8_t<9>(P, S, slice, a);
  else sum_slices8_t<16>(P, S, slice, a);
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ P, int S, int M, int N,
                                                            bf16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x;
  const long long slice = (long long)M * N;
  for (int v = threadIdx.x; v < (N >> 3); v += blockDim.x) {
    float a[8];
    sum_slices8(P + row * N + v * 8, S, slice, a);
    *reinterpret_cast<uint4*>(out + row * ldo + v * 8) =
        make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
  }
}

template <int kVecPerThread>
__global__ void __launch_bounds__(256) splitk_reduce_rmsnorm_kernel(const float* __restrict__ P, int S, int M, int N,
                                                                    bf16* __restrict__ x, long long ldx,
                                                                    const bf16* __restrict__ w, bf16* __restrict__ xn,
                                                                    long long ldxn, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[33];
  const long long row = blockIdx.x;
  const long long slice = (long long)M * N;
  const int nvec = N >> 3;
  uint32_t keep[kVecPerThread][4];  // the new residual row, packed bf16
  uint4 wv[kVecPerThread];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec) {
      const uint4 r = *reinterpret_cast<const uint4*>(x + row * ldx + v * 8);  // requested before the slice sums wait
      wv[i] = *reinterpret_cast<const uint4*>(w + v * 8);
      float a[8];
      sum_slices8(P + row * N + v * 8, S, slice, a);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = bf16_round(bf16lo(rr[j]) + bf16_round(a[2 * j]));
        const float hi = bf16_round(bf16hi(rr[j]) + bf16_round(a[2 * j + 1]));
        ss = fmaf(lo, lo, ss);
        ss = fmaf(hi, hi, ss);
        keep[i][j] = pack_bf16(lo, hi);
      }
      *reinterpret_cast<uint4*>(x + row * ldx + v * 8) = make_uint4(keep[i][0], keep[i][1], keep[i][2], keep[i][3]);
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / (float)N + eps);
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec) {
      const uint32_t b[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = bf16_round(bf16lo(keep[i][j]) * inv) * bf16lo(b[j]);
        const float hi = bf16_round(bf16hi(keep[i][j]) * inv) * bf16hi(b[j]);
        o[j] = pack_bf16(lo, hi);
      }
      *reinterpret_cast<uint4*>(xn + row * ldxn + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Per-head q/k RMSNorm + RoPE + KV-cache append (HF Qwen3Attention.forward, installed
// modeling_qwen3.py:236-271; apply_rotary_pos_emb :86-181; DynamicCache.update replaced by an in-place
// paged store). One warp per (row, head); head_dim == 128: lane l owns elements {2l, 2l+1} and their
// rotate-half partners {64+2l, 65+2l}.
//   qkv   [rows, (Hq + 2*Hkv) * 128] bf16 (output of the fused q/k/v projection)
//   q_out [rows, Hq * 128] bf16
//   k_pool/v_pool [num_pages, Hkv, page_size, 128] bf16; page = block_table[seq * max_pages + pos / page_size]
//                 (block_table == NULL: page = seq * max_pages + pos / page_size, i.e. a contiguous cache)
// cos/sin are evaluated in fp32 from pos * inv_freq and rounded to bf16 before use, products and the sum
// are rounded to bf16 one by one, as the bf16 eager reference does.
// ------------------------------------------------------------------------------------------------
struct RopeParams {
  const bf16* qkv;
  long long ld_qkv;
  const bf16* q_norm_w;
  const bf16* k_norm_w;
  const float* inv_freq;  // [64]
  const int* positions;   // [rows]
  const int* row_seq;     // [rows] sequence index of each row (NULL: row index)
  bf16* q_out;
  bf16* k_pool;
  bf16* v_pool;
  const int* block_table;
  int max_pages, page_shift, num_pages;
  int rows, Hq, Hkv;
  float eps;
  int* err_flag;
};

// One CTA (8 warps) per ROW: the 64 angles of the row are evaluated once (the accurate cosf / sinf with arguments up to
// ~12 000 rad were most of this kernel's time when every 8-head group of a row recomputed them), every warp takes heads
// warp, warp + 8, ... and requests the loads of all its heads before it touches the first one.
__global__ void __launch_bounds__(256) qknorm_rope_kv_kernel(const RopeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_cs[128];  // cos[64] | sin[64] of this row, bf16-rounded
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int heads = p.Hq + 2 * p.Hkv;
  const int pos = p.positions[row];
  if (threadIdx.x < 64) {
    const float f = (float)pos * p.inv_freq[threadIdx.x];
    s_cs[threadIdx.x] = bf16_round(cosf(f));
    s_cs[64 + threadIdx.x] = bf16_round(sinf(f));
  }
  constexpr int kMaxPerWarp = 4;  // 32 heads over 8 warps; more heads run in further rounds
  const bf16* src_row = p.qkv + (long long)row * p.ld_qkv;
  const int seq = p.row_seq ? p.row_seq[row] : row;
  const uint32_t wqa = *reinterpret_cast<const uint32_t*>(p.q_norm_w + 2 * lane), wqb = *reinterpret_cast<const uint32_t*>(p.q_norm_w + 64 + 2 * lane);
  const uint32_t wka = *reinterpret_cast<const uint32_t*>(p.k_norm_w + 2 * lane), wkb = *reinterpret_cast<const uint32_t*>(p.k_norm_w + 64 + 2 * lane);
  __syncthreads();
  const float c0 = s_cs[2 * lane], c1 = s_cs[2 * lane + 1], s0 = s_cs[64 + 2 * lane], s1 = s_cs[65 + 2 * lane];
  for (int h0 = warp; h0 < heads; h0 += 8 * kMaxPerWarp) {
    uint32_t a[kMaxPerWarp], b[kMaxPerWarp];
#pragma unroll
    for (int i = 0; i < kMaxPerWarp; ++i) {
      const int h = h0 + 8 * i;
      if (h < heads) {
        a[i] = *reinterpret_cast<const uint32_t*>(src_row + h * 128 + 2 * lane);       // elements 2l, 2l+1
        b[i] = *reinterpret_cast<const uint32_t*>(src_row + h * 128 + 64 + 2 * lane);  // elements 64+2l, 65+2l
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxPerWarp; ++i) {
      const int h = h0 + 8 * i;
      if (h >= heads) break;
      bf16* dst;
      if (h < p.Hq) {
        dst = p.q_out + ((long long)row * p.Hq + h) * 128;
      } else {
        const int hk = (h - p.Hq) % p.Hkv;
        const bool is_v = (h - p.Hq) >= p.Hkv;
        const int lp = pos >> p.page_shift;
        int page = -1;
        if (pos >= 0 && lp < p.max_pages) page = p.block_table ? p.block_table[(long long)seq * p.max_pages + lp] : seq * p.max_pages + lp;
        if (page < 0 || page >= p.num_pages) {
          if (p.err_flag && lane == 0) *p.err_flag = 2;
          continue;
        }
        bf16* pool = is_v ? p.v_pool : p.k_pool;
        const int slot = pos & ((1 << p.page_shift) - 1);
        dst = pool + (((long long)page * p.Hkv + hk) << p.page_shift) * 128 + (long long)slot * 128;
        if (is_v) {  // V: plain copy
          *reinterpret_cast<uint32_t*>(dst + 2 * lane) = a[i];
          *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = b[i];
          continue;
        }
      }
      const uint32_t wa = (h < p.Hq) ? wqa : wka, wb = (h < p.Hq) ? wqb : wkb;
      float x0 = bf16lo(a[i]), x1 = bf16hi(a[i]), x2 = bf16lo(b[i]), x3 = bf16hi(b[i]);
      float ss = x0 * x0;
      ss = fmaf(x1, x1, ss);
      ss = fmaf(x2, x2, ss);
      ss = fmaf(x3, x3, ss);
      ss = warp_sum(ss);
      const float inv = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
      // normed = w * bf16(x * inv), rounded to bf16 (it is a bf16 tensor in the reference)
      x0 = bf16_round(bf16lo(wa) * bf16_round(x0 * inv));
      x1 = bf16_round(bf16hi(wa) * bf16_round(x1 * inv));
      x2 = bf16_round(bf16lo(wb) * bf16_round(x2 * inv));
      x3 = bf16_round(bf16hi(wb) * bf16_round(x3 * inv));
      // RoPE: out[i] = bf16(x[i]*cos) + bf16(-x[i+64]*sin) ; out[i+64] = bf16(x[i+64]*cos) + bf16(x[i]*sin)
      const float o0 = bf16_round(x0 * c0) + bf16_round(-x2 * s0);
      const float o1 = bf16_round(x1 * c1) + bf16_round(-x3 * s1);
      const float o2 = bf16_round(x2 * c0) + bf16_round(x0 * s0);
      const float o3 = bf16_round(x3 * c1) + bf16_round(x1 * s1);
      *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pack_bf16(o0, o1);
      *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = pack_bf16(o2, o3);
    }
  }
}


// Prompt-sized variant (heads % 4 == 0, heads <= 32): the same arithmetic bit for bit, with 16-byte accesses. Eight lanes
// share a head: lane j of the group holds elements 8j..8j+7 and 64+8j..64+8j+7, i.e. the data of the four lanes
// 4j..4j+3 of the kernel above ("virtual lanes"), so the sum of squares is rebuilt in exactly that kernel's order: the
// per-lane fma chain over (2v, 2v+1, 64+2v, 65+2v), the butterfly levels xor 16 / 8 / 4 as shuffles over lanes
// xor 4 / 2 / 1, and the levels xor 2 / 1 inside the lane. A warp covers 4 heads, the 8 warps one row; a CTA walks
// kWideRows consecutive rows: their 4 x 64 angles are evaluated by the 256 threads at once (one barrier per CTA instead
// of one per row, a quarter of the CTA launches) and the loads of all four rows are requested before the first is used.
// 2.2 GB per launch at a bench-sized prefill; the 4-byte kernel ran at 3.1 TB/s there.
constexpr int kWideRows = 4;
__global__ void __launch_bounds__(256, 3) qknorm_rope_kv_wide_kernel(const RopeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float s_cs[kWideRows][128];  // cos[64] | sin[64] per row, bf16-rounded
  const int row0 = blockIdx.x * kWideRows;
  const int nrows = min(kWideRows, p.rows - row0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, j = lane & 7;
  const int heads = p.Hq + 2 * p.Hkv;
  const int h = warp * 4 + sub;          // this lane group's head; heads % 4 == 0 makes `h < heads` uniform per warp
  const bool active = h < heads;
  const long long col = (long long)h * 128 + 8 * j;
  uint4 ra[kWideRows], rb[kWideRows];  // all rows of the CTA requested up front: 32 KB in flight per CTA
#pragma unroll
  for (int r = 0; r < kWideRows; ++r) {
    ra[r] = make_uint4(0, 0, 0, 0);
    rb[r] = ra[r];
    if (active && r < nrows) {
      const bf16* src = p.qkv + (long long)(row0 + r) * p.ld_qkv + col;
      ra[r] = *reinterpret_cast<const uint4*>(src);
      rb[r] = *reinterpret_cast<const uint4*>(src + 64);
    }
  }
  {
    const int r = threadIdx.x >> 6, i = threadIdx.x & 63;
    if (r < nrows) {
      const float f = (float)p.positions[row0 + r] * p.inv_freq[i];
      s_cs[r][i] = bf16_round(cosf(f));
      s_cs[r][64 + i] = bf16_round(sinf(f));
    }
  }
  __syncthreads();
  if (!active) return;
#pragma unroll
  for (int r = 0; r < kWideRows; ++r) {
    if (r >= nrows) break;
    const int row = row0 + r;
    const uint4 a = ra[r], b = rb[r];
    const int pos = p.positions[row];
    bf16* dst;
    bool is_v = false, ok = true;
    if (h < p.Hq) {
      dst = p.q_out + ((long long)row * p.Hq + h) * 128;
    } else {
      const int seq = p.row_seq ? p.row_seq[row] : row;
      const int hk = (h - p.Hq) % p.Hkv;
      is_v = (h - p.Hq) >= p.Hkv;
      const int lp = pos >> p.page_shift;
      int page = -1;
      if (pos >= 0 && lp < p.max_pages) page = p.block_table ? p.block_table[(long long)seq * p.max_pages + lp] : seq * p.max_pages + lp;
      if (page < 0 || page >= p.num_pages) {
        if (p.err_flag && j == 0) *p.err_flag = 2;
        ok = false;
        page = 0;
      }
      bf16* pool = is_v ? p.v_pool : p.k_pool;
      const int slot = pos & ((1 << p.page_shift) - 1);
      dst = pool + (((long long)page * p.Hkv + hk) << p.page_shift) * 128 + (long long)slot * 128;
    }
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
    float s3[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float x0 = bf16lo(ua[m]), x1 = bf16hi(ua[m]), x2 = bf16lo(ub[m]), x3 = bf16hi(ub[m]);
      float ss = x0 * x0;
      ss = fmaf(x1, x1, ss);
      ss = fmaf(x2, x2, ss);
      ss = fmaf(x3, x3, ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);  // virtual lanes v ^ 16
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);  // v ^ 8
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);  // v ^ 4
      s3[m] = ss;
    }
    if (ok) {
      if (is_v) {  // V: plain copy
        *reinterpret_cast<uint4*>(dst + 8 * j) = a;
        *reinterpret_cast<uint4*>(dst + 64 + 8 * j) = b;
      } else {
        const float ssum = (s3[0] + s3[2]) + (s3[1] + s3[3]);  // v ^ 2, then v ^ 1
        const float inv = rsqrtf(ssum * (1.0f / 128.0f) + p.eps);
        const bf16* nw = (h < p.Hq) ? p.q_norm_w : p.k_norm_w;
        const uint4 w0 = *reinterpret_cast<const uint4*>(nw + 8 * j), w1 = *reinterpret_cast<const uint4*>(nw + 64 + 8 * j);
        const uint32_t wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
        uint32_t oa[4], ob[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          float x0 = bf16lo(ua[m]), x1 = bf16hi(ua[m]), x2 = bf16lo(ub[m]), x3 = bf16hi(ub[m]);
          x0 = bf16_round(bf16lo(wa[m]) * bf16_round(x0 * inv));
          x1 = bf16_round(bf16hi(wa[m]) * bf16_round(x1 * inv));
          x2 = bf16_round(bf16lo(wb[m]) * bf16_round(x2 * inv));
          x3 = bf16_round(bf16hi(wb[m]) * bf16_round(x3 * inv));
          const float2 cc = *reinterpret_cast<const float2*>(&s_cs[r][8 * j + 2 * m]);
          const float2 tt = *reinterpret_cast<const float2*>(&s_cs[r][64 + 8 * j + 2 * m]);
          const float c0 = cc.x, c1 = cc.y, s0 = tt.x, s1 = tt.y;
          const float o0 = bf16_round(x0 * c0) + bf16_round(-x2 * s0);
          const float o1 = bf16_round(x1 * c1) + bf16_round(-x3 * s1);
          const float o2 = bf16_round(x2 * c0) + bf16_round(x0 * s0);
          const float o3 = bf16_round(x3 * c1) + bf16_round(x1 * s1);
          oa[m] = pack_bf16(o0, o1);
          ob[m] = pack_bf16(o2, o3);
        }
        *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<uint4*>(dst + 64 + 8 * j) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
      }
    }
  }
}

}  // namespace

extern "C" int mtts_embed_sum8(const long long* ids, int rows, int channels, const void* const* tables_host,
                               const int* vocab_sizes_host, int hidden, void* out, int* err_flag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(channels >= 1 && channels <= 8, "mtts_embed_sum8: channels must be in [1,8], got %d", channels);
  MTTS_REQUIRE(hidden > 0 && hidden % 8 == 0, "mtts_embed_sum8: hidden must be a multiple of 8");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(ids && tables_host && vocab_sizes_host && out, "mtts_embed_sum8: null pointer");
  EmbedParams p;
  p.ids = ids;
  for (int c = 0; c < 8; ++c) {
    p.tables[c] = c < channels ? reinterpret_cast<const bf16*>(tables_host[c]) : nullptr;
    p.vocab[c] = c < channels ? vocab_sizes_host[c] : 0;
  }
  p.rows = rows; p.channels = channels; p.hidden = hidden;
  p.out = reinterpret_cast<bf16*>(out);
  p.err_flag = err_flag;
  const long long total = (long long)rows * (hidden / 8);
  MTTS_CUDA_CHECK(mtts_launch(embed_sum_kernel, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, stream, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_rmsnorm(const void* x, long long ldx, const void* w, void* out, long long ldo, int rows, int hidden,
                            float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(hidden > 0 && hidden % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0,
               "mtts_rmsnorm: hidden and strides must be multiples of 8");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(x && w && out, "mtts_rmsnorm: null pointer");
  if (hidden == 2048) {  // a warp per row, the row in registers: 2 rows per CTA at decode sizes (spread over the SMs),
    const int wpc = rows <= 1024 ? 2 : 8;  // 8 at prompt sizes (32 KB of loads in flight per CTA, one pass over the data)
    MTTS_CUDA_CHECK(mtts_launch(rmsnorm_warp_kernel, dim3((rows + wpc - 1) / wpc), dim3(32 * wpc), 0, stream,
                                reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(w),
                                reinterpret_cast<bf16*>(out), ldo, rows, eps));
    MTTS_LAUNCH_CHECK();
    return MTTS_OK;
  }
  MTTS_CUDA_CHECK(mtts_launch(rmsnorm_kernel, dim3(rows), dim3(256), 0, stream, reinterpret_cast<const bf16*>(x), ldx,
                              reinterpret_cast<const bf16*>(w), reinterpret_cast<bf16*>(out), ldo, hidden, eps));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_splitk_reduce(const float* partials, int splits, int M, int N, void* out, long long ldo, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(splits >= 1 && splits <= 16 && N % 8 == 0 && ldo % 8 == 0, "mtts_splitk_reduce: N and ldo must be multiples of 8, splits <= 16");
  if (M <= 0) return MTTS_OK;
  MTTS_REQUIRE(partials && out, "mtts_splitk_reduce: null pointer");
  MTTS_CUDA_CHECK(mtts_launch(splitk_reduce_kernel, dim3(M), dim3(256), 0, stream, partials, splits, M, N,
                              reinterpret_cast<bf16*>(out), ldo));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_splitk_reduce_rmsnorm(const float* partials, int splits, int M, int N, void* x, long long ldx,
                                          const void* norm_w, void* xn, long long ldxn, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(splits >= 1 && splits <= 16 && N % 8 == 0 && N <= 8192 && ldx % 8 == 0 && ldxn % 8 == 0,
               "mtts_splitk_reduce_rmsnorm: N (<= 8192) and strides must be multiples of 8");
  if (M <= 0) return MTTS_OK;
  MTTS_REQUIRE(partials && x && norm_w && xn, "mtts_splitk_reduce_rmsnorm: null pointer");
  const bf16* w = reinterpret_cast<const bf16*>(norm_w);
  bf16* xp = reinterpret_cast<bf16*>(x);
  bf16* xnp = reinterpret_cast<bf16*>(xn);
  if (N <= 2048)
    MTTS_CUDA_CHECK(mtts_launch(splitk_reduce_rmsnorm_kernel<1>, dim3(M), dim3(256), 0, stream, partials, splits, M, N, xp, ldx,
                                w, xnp, ldxn, eps));
  else if (N <= 4096)
    MTTS_CUDA_CHECK(mtts_launch(splitk_reduce_rmsnorm_kernel<2>, dim3(M), dim3(256), 0, stream, partials, splits, M, N, xp, ldx,
                                w, xnp, ldxn, eps));
  else
    MTTS_CUDA_CHECK(mtts_launch(splitk_reduce_rmsnorm_kernel<4>, dim3(M), dim3(256), 0, stream, partials, splits, M, N, xp, ldx,
                                w, xnp, ldxn, eps));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_qknorm_rope_kvappend(const void* qkv, long long ld_qkv, const void* q_norm_w, const void* k_norm_w,
                                         const float* inv_freq, const int* positions, const int* row_seq, void* q_out,
                                         void* k_pool, void* v_pool, const int* block_table, int max_pages,
                                         int page_size, int num_pages, int rows, int num_q_heads, int num_kv_heads,
                                         int head_dim, float eps, int* err_flag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == 128, "mtts_qknorm_rope_kvappend: head_dim must be 128 (got %d)", head_dim);
  MTTS_REQUIRE(page_size > 0 && (page_size & (page_size - 1)) == 0, "mtts_qknorm_rope_kvappend: page_size must be a power of two");
  if (rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(qkv && q_norm_w && k_norm_w && inv_freq && positions && q_out && k_pool && v_pool,
               "mtts_qknorm_rope_kvappend: null pointer");
  RopeParams p;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.ld_qkv = ld_qkv;
  p.q_norm_w = reinterpret_cast<const bf16*>(q_norm_w); p.k_norm_w = reinterpret_cast<const bf16*>(k_norm_w);
  p.inv_freq = inv_freq; p.positions = positions; p.row_seq = row_seq;
  p.q_out = reinterpret_cast<bf16*>(q_out); p.k_pool = reinterpret_cast<bf16*>(k_pool);
  p.v_pool = reinterpret_cast<bf16*>(v_pool); p.block_table = block_table; p.max_pages = max_pages;
  int shift = 0;
  while ((1 << shift) < page_size) ++shift;
  p.page_shift = shift; p.num_pages = num_pages; p.rows = rows; p.Hq = num_q_heads; p.Hkv = num_kv_heads; p.eps = eps;
  p.err_flag = err_flag;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int heads = num_q_heads + 2 * num_kv_heads;
  const bool wide = heads % 4 == 0 && heads <= 32 && ld_qkv % 8 == 0 && al16(qkv) && al16(q_out) && al16(k_pool) &&
                    al16(v_pool) && al16(q_norm_w) && al16(k_norm_w);
  if (wide) {  // 16-byte accesses, same bits (see the kernel)
    MTTS_CUDA_CHECK(mtts_launch(qknorm_rope_kv_wide_kernel, dim3((unsigned)((rows + kWideRows - 1) / kWideRows)), dim3(256), 0,
                                stream, p));
  } else {
    MTTS_CUDA_CHECK(mtts_launch(qknorm_rope_kv_kernel, dim3((unsigned)rows), dim3(256), 0, stream, p));
  }
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
