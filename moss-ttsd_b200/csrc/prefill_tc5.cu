// Causal GQA prefill attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 128, bf16: the attention of
// HF Qwen3Attention.forward (installed modeling_qwen3.py:236-288, eager_attention_forward) for the packed prompt rows of a
// prefill, invoked from AsteroidTTSInstruct.forward (modeling_asteroid.py:226,273-284). Replaces the mma.sync flash kernel
// (gqa_prefill_fa_kernel, attention.cu) behind mtts_gqa_attention(rows_per_tile = 128).
//
//   q   [rows, Hq * 128] bf16   (q-norm + RoPE applied, mtts_qknorm_rope_kvappend)
//   K/V pools [num_pages, Hkv, page_size, 128] bf16: the keys of a sequence incl. the rows of this prefill are already
//       in the pool; a 64-key tile of one (page, kv head) is 64 consecutive rows of the [pages * Hkv * page_size, 128] view
//   out [rows, Hq * 128] bf16
//
// One CTA (128 threads) owns up to 128 consecutive query rows of ONE sequence and one q head, and walks keys 0 .. last
// position of the tile in tiles of 64 (same structure as mha_tc5.cu):
//   S = Q K^T   tcgen05.mma M=128 N=64 K=128 (two 64-dim halves), S in TMEM (64 columns)
//   softmax     thread = query row (tcgen05.ld 32x32b), causal mask only on tiles that reach the diagonal, exp2 with the scale
//               folded in, P in bf16 (as the reference casts the probabilities) to shared memory, 128B-swizzled K-major
//   O += P V    tcgen05.mma M=128 N=128 K=64, V tile MN-major as TMA stored it (two 64-dim blocks, LBO apart), O in TMEM
//               (128 columns), rescaled in place only when a row of the warp saw a new maximum
// K and V in two-stage rings (a K tile is free once S is computed, a V tile once O += P V has retired); warps 0..3 are the
// softmax threads, warp 4 issues every TMA copy and MMA, the two sides meet on mbarriers only (no block barrier in the key
// loop). 112 KB of shared memory and 256 TMEM columns per CTA, two CTAs per SM.
#include "common.cuh"
#include "sm100.cuh"
#include "mtts_internal.h"

using namespace sm100;

namespace {

constexpr int kQ = 128, kK = 64, kD = 128;
constexpr uint32_t kSub = 64 * 128;  // 8 KB: 64 rows x 64 dims (128 B), one TMA box
// Q 32 KB + K ring 32 KB + V ring 32 KB + P 16 KB + barriers. No alignment slack: two CTAs must fit one SM's 228 KB
// (2 x (112 KB + 128 B + 1 KB reserved) = 226.25 KB), so the kernel checks that its dynamic shared memory starts 1024-aligned.
constexpr uint32_t kSmemBytes = 4 * kSub /*Q*/ + 2 * 2 * kSub /*K ring*/ + 2 * 2 * kSub /*V ring*/ + 2 * kSub /*P*/ + 128;
constexpr int kThreadsPf = 160;  // warps 0..3: softmax (thread = query row), warp 4: control (TMA + MMA issue)

struct PrefillParams {
  bf16* out;
  const int* block_table;
  const int* tile_row0;
  const int* tile_nrows;
  const int* row_seq;
  const int* positions;
  int max_pages, page_shift, Hq, Hkv, rows;
  float scale_log2;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: 64-element (128 B) column blocks `kSub` bytes apart (LBO), groups of 8 rows of the
// other dimension 1024 B apart (SBO)
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kSub >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreadsPf, 2) gqa_prefill_tc5_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                        const __grid_constant__ CUtensorMap tm_k,
                                                                        const __grid_constant__ CUtensorMap tm_v, const PrefillParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                 // [2 dim halves][128 q][64 d]
  uint8_t* sK = sQ + 4 * kSub;        // [2 stages][2 dim halves][64 keys][64 d]
  uint8_t* sV = sK + 4 * kSub;        // [2 stages][2 dim halves][64 keys][64 d]
  uint8_t* sP = sV + 4 * kSub;        // [128 q][64 keys]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sP + 2 * kSub);
  uint64_t* k_full = q_full + 1;      // [2]
  uint64_t* v_full = k_full + 2;      // [2]
  uint64_t* s_full = v_full + 2;
  uint64_t* pv_done = s_full + 1;
  uint64_t* s_free = pv_done + 1;     // 4 arrivals: every softmax warp holds its scores of the current tile
  uint64_t* p_ready = s_free + 1;     // 4 arrivals: every softmax warp has written its rows of P (and rescaled O)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(p_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, hq = blockIdx.y;
  const int hk = hq / (p.Hq / p.Hkv);
  const int row0 = p.tile_row0[tile], nrows = p.tile_nrows[tile];
  const int seq = p.row_seq ? p.row_seq[row0] : tile;
  const int pos0 = p.positions[row0];            // rows of a tile are consecutive positions of one sequence
  const int n_tiles = (pos0 + nrows - 1) / kK + 1;
  const int page_mask = (1 << p.page_shift) - 1;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("mtts: prefill attention: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_k);
    prefetch_tmap(&tm_v);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1);
    mbar_init(&k_full[1], 1);
    mbar_init(&v_full[0], 1);
    mbar_init(&v_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(pv_done, 1);
    mbar_init(s_free, 4);
    mbar_init(p_ready, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
    __syncwarp();
    tmem_alloc<256>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  constexpr uint32_t kIdescS = make_idesc(1, kQ, kK);               // bf16 x bf16 -> f32, both K-major
  constexpr uint32_t kIdescO = make_idesc(1, kQ, kD) | (1u << 16);  // B (the V tile) MN-major

  if (warp == 4) {
    // ================= control: TMA copies + MMA issue =================
    if (lane == 0) {
      // first row of key tile t in the [pages * Hkv * page_size, 128] view of the pools
      auto kv_row = [&](int t) -> int {
        const int key0 = t * kK;
        const int lp = key0 >> p.page_shift;
        const int page = p.block_table ? __ldg(p.block_table + (long long)seq * p.max_pages + lp) : seq * p.max_pages + lp;
        return ((page * p.Hkv + hk) << p.page_shift) + (key0 & page_mask);
      };
      auto load_k = [&](int t) {
        const int r = kv_row(t), st = t & 1;
        mbar_arrive_expect_tx(&k_full[st], 2 * kSub);
        tma_load_2d(sK + st * 2 * kSub, &tm_k, &k_full[st], 0, r, kEvictLast);
        tma_load_2d(sK + st * 2 * kSub + kSub, &tm_k, &k_full[st], 64, r, kEvictLast);
      };
      auto load_v = [&](int t) {
        const int r = kv_row(t), st = t & 1;
        mbar_arrive_expect_tx(&v_full[st], 2 * kSub);
        tma_load_2d(sV + st * 2 * kSub, &tm_v, &v_full[st], 0, r, kEvictLast);
        tma_load_2d(sV + st * 2 * kSub + kSub, &tm_v, &v_full[st], 64, r, kEvictLast);
      };
      auto issue_s = [&](int t) {
        mbar_wait(&k_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + (t & 1) * 2 * kSub);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t half = k >> 2, in = (k & 3) * 32;
          umma_bf16(tmem_base, make_smem_desc_sw128(qa + half * 2 * kSub + in), make_smem_desc_sw128(ka + half * kSub + in), kIdescS,
                    k > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      mbar_arrive_expect_tx(q_full, 4 * kSub);
      const int cq = hq * kD;
      tma_load_2d(sQ, &tm_q, q_full, cq, row0, kEvictNormal);
      tma_load_2d(sQ + kSub, &tm_q, q_full, cq, row0 + 64, kEvictNormal);
      tma_load_2d(sQ + 2 * kSub, &tm_q, q_full, cq + 64, row0, kEvictNormal);
      tma_load_2d(sQ + 3 * kSub, &tm_q, q_full, cq + 64, row0 + 64, kEvictNormal);
      for (int t = 0; t < 2 && t < n_tiles; ++t) {
        load_k(t);
        load_v(t);
      }
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(s_free, j & 1);  // S_j is in registers: its TMEM columns and K stage j % 2 are free
        tc_fence_after();
        if (j + 2 < n_tiles) load_k(j + 2);
        if (j + 1 < n_tiles) issue_s(j + 1);  // runs under tile j's softmax
        mbar_wait(p_ready, j & 1);            // (the softmax warps waited for v_full themselves on the last tile)
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (j & 1) * 2 * kSub);
#pragma unroll
        for (int k = 0; k < kK / 16; ++k)
          umma_bf16(tmem_base + 64, make_smem_desc_sw128(pa + k * 32), make_smem_desc_mn_sw128(va + k * 2048), kIdescO,
                    (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(pv_done);
        if (j + 2 < n_tiles) {  // V stage j % 2 is free once this O += P V has retired
          mbar_wait(pv_done, j & 1);
          load_v(j + 2);
        }
      }
    }
  } else {
    // ================= softmax: thread = query row =================
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);  // columns 0..63: S
    const uint32_t tO = tS + 64;                                               // columns 64..191: O
    float m_run = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2;
    const int row = warp * 32 + lane;
    const int my_pos = pos0 + row;  // keys <= my_pos are visible to this row
    uint8_t* p_row = sP + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(tS, s0);
      tmem_ld_32x32b_x32(tS + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      // ---- causal mask (only tiles that reach this row's diagonal), row maximum, probabilities
      const int lim = my_pos - j * kK;  // keys 0..lim of this tile are visible (lim >= 63: all)
      float mx = -INFINITY;
      if (lim < kK - 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i > lim) s0[i] = __float_as_uint(-INFINITY);
          if (32 + i > lim) s1[i] = __float_as_uint(-INFINITY);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      float m_new = fmaxf(m_run, mx);
      // a row with nothing visible yet can only be a row beyond the tile's sequence (never stored): keep it finite
      if (m_new == -INFINITY) m_new = 0.f;
      // LAZY rescaling: the reference maximum of a row moves only when the true maximum has outgrown it by more than 2^8
      // (in the exp2 domain) somewhere in the warp; until then probabilities may exceed 1 (<= 256: harmless in bf16 / fp32)
      // and O, l keep their scale — the result is the same quotient. Only rows of this tile's sequence vote: a row's
      // arithmetic must not depend on which other sequence follows it in the packed batch.
      const bool move = j == 0 || __any_sync(0xffffffffu, row < nrows && (m_new - m_run) * c > 8.0f);
      const float m_use = move ? m_new : m_run;
      const float alpha = ex2((m_run - m_use) * c);
      const float mc = m_use * c;
      float sum = 0.f;
      uint32_t ph[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = ex2(fmaf(__uint_as_float(s0[2 * i]), c, -mc)), a1 = ex2(fmaf(__uint_as_float(s0[2 * i + 1]), c, -mc));
        const float b0 = ex2(fmaf(__uint_as_float(s1[2 * i]), c, -mc)), b1 = ex2(fmaf(__uint_as_float(s1[2 * i + 1]), c, -mc));
        const uint32_t pa = pack_bf16(a0, a1), pb = pack_bf16(b0, b1);
        sum += (bf16lo(pa) + bf16hi(pa)) + (bf16lo(pb) + bf16hi(pb));  // the sum runs over the rounded probabilities
        ph[i] = pa;
        ph[16 + i] = pb;
      }
      l_run = l_run * alpha + sum;
      m_run = m_use;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // O += P V of the previous tile retired: P and O are free
        tc_fence_after();
        if (move) {  // warp-uniform
          uint32_t o[32];
#pragma unroll 1
          for (int q4 = 0; q4 < 4; ++q4) {
            tmem_ld_32x32b_x32(tO + q4 * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tO + q4 * 32, o);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(p_row + ((ch ^ sw) << 4)) = make_uint4(ph[4 * ch], ph[4 * ch + 1], ph[4 * ch + 2], ph[4 * ch + 3]);
      {
        // keys past the last row of this prefill are not in the pool yet: their probabilities are 0, but 0 x stale NaN / Inf
        // bits of V would still poison O — zero those V rows (a row's 128 bytes stay inside the row under the swizzle)
        const int nvalid = pos0 + nrows - j * kK;
        if (nvalid < kK) {  // uniform: last key tile only
          mbar_wait(&v_full[j & 1], (j >> 1) & 1);
          uint8_t* vb = sV + (j & 1) * 2 * kSub;
          for (int i = tid; i < (kK - nvalid) * 16; i += 128) {
            const int kr = nvalid + (i >> 4), part = i & 15;
            *reinterpret_cast<uint4*>(vb + (part >> 3) * kSub + kr * 128 + (part & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    bf16* op = p.out + ((long long)(row0 + row)) * (p.Hq * kD) + hq * kD;
#pragma unroll 1
    for (int q4 = 0; q4 < 4; ++q4) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + q4 * 32, o);
      tmem_ld_wait();
      if (row < nrows) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16(__uint_as_float(o[v4 * 8 + 2 * i]) * inv, __uint_as_float(o[v4 * 8 + 2 * i + 1]) * inv);
          *reinterpret_cast<uint4*>(op + q4 * 32 + v4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

int mtts_configure_prefill_tc5() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gqa_prefill_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return MTTS_OK;
}

extern "C" int mtts_gqa_prefill_tc(const void* q, long long rows, const void* k_pool, const void* v_pool, const int* block_table,
                                   int max_pages, int page_size, int num_pages, const int* tile_row0, const int* tile_nrows,
                                   const int* row_seq, const int* positions, void* out, int tiles, int Hq, int Hkv, int head_dim,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kD, "mtts_gqa_prefill_tc: head_dim must be 128 (got %d)", head_dim);
  MTTS_REQUIRE(Hkv > 0 && Hq % Hkv == 0, "mtts_gqa_prefill_tc: bad head counts");
  MTTS_REQUIRE(page_size >= kK && (page_size & (page_size - 1)) == 0, "mtts_gqa_prefill_tc: page_size must be a power of two >= 64");
  if (tiles <= 0 || rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(q && k_pool && v_pool && positions && out && tile_row0 && tile_nrows && num_pages > 0, "mtts_gqa_prefill_tc: null pointer");
  int page_shift = 0;
  while ((1 << page_shift) < page_size) ++page_shift;
  const long long pool_rows = ((long long)num_pages * Hkv) << page_shift;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  CUtensorMap tq, tk, tv;
  int rc = mtts_get_tmap_2d(q, rows, (long long)Hq * kD, (long long)Hq * kD, 64, 2, &tq);
  if (rc) return rc;
  rc = mtts_get_tmap_2d(k_pool, pool_rows, kD, kD, 64, 2, &tk);
  if (rc) return rc;
  rc = mtts_get_tmap_2d(v_pool, pool_rows, kD, kD, 64, 2, &tv);
  if (rc) return rc;
  PrefillParams p;
  p.out = reinterpret_cast<bf16*>(out);
  p.block_table = block_table; p.tile_row0 = tile_row0; p.tile_nrows = tile_nrows; p.row_seq = row_seq; p.positions = positions;
  p.max_pages = max_pages; p.page_shift = page_shift; p.Hq = Hq; p.Hkv = Hkv; p.rows = (int)rows;
  p.scale_log2 = scale_log2;
  MTTS_CUDA_CHECK(mtts_launch(gqa_prefill_tc5_kernel, dim3(tiles, Hq), dim3(kThreadsPf), kSmemBytes, stream, tq, tk, tv, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
