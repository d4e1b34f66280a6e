// Dense projection GEMM on the 5th-gen tensor cores (tcgen05 + TMEM accumulators + TMA-fed operands).
//
//   out[m, n] = epilogue( sum_k x[m, k] * w[n, k] )        x: [M, K] row-major, w: [N, K] row-major
//
// This is the "nn.Linear" contraction used by every dense projection on the hot path
// (reference call sites: HF Qwen3Attention/Qwen3MLP q/k/v/o/gate/up/down invoked from
// modeling_asteroid.py:226,273-284; the 8 lm_heads modeling_asteroid.py:412; codec nn.Linear /
// ConvTranspose1d-as-GEMM XY_Tokenizer/xy_tokenizer/nn/modules.py:84-87,181-182,494-500,1111-1115).
//
// B200 design (not a translation of anything in the reference, which only calls aten::linear):
//   * swap-AB: the WEIGHT tile (128 rows of w) is the UMMA "A"/M operand and the activation tile
//     (BN rows of x, BN in {16..256}) is the UMMA "B"/N operand, so a batch-1 decode step still
//     issues full 128-lane MMAs and the kernel degenerates to a pure weight-streaming pass that is
//     HBM-bound; large M (prefill, codec) uses BN = 128/256.
//   * operands arrive by TMA (cp.async.bulk.tensor, 128B swizzle) into a kStages-deep mbarrier ring;
//     one elected thread issues tcgen05.mma; accumulators live in TMEM; four epilogue warps read them
//     back with tcgen05.ld, stage the tile (shared memory, or the split-K workspace) and run a row-major,
//     16-byte-vectorised epilogue (bias / GELU / layer-scale / residual / SwiGLU / the reference's bf16
//     rounding points).
//   * split-K (grid.z) keeps all SMs pulling weights when N/128 tiles are few. The splits of one output tile
//     form a thread-block CLUSTER: every CTA stages its partial tile in its own shared memory, and after one
//     cluster barrier each CTA reduces a slice of the rows by reading its peers' tiles over distributed shared
//     memory in fixed rank order (bitwise deterministic; no workspace, no atomics, no global round trips).
//   * PDL: the kernel starts while its predecessor in the stream is still running and streams its WEIGHT
//     tiles (which never depend on the predecessor) into the ring before griddepcontrol.wait; activations,
//     residual reads and all global writes happen after it. The ring is sized so that two CTAs (this GEMM's
//     and the next one's) fit on an SM.
#include "common.cuh"
#include "sm100.cuh"
#include "mtts_internal.h"

#include <cuda_fp16.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include <stdlib.h>

using namespace sm100;

namespace {

#ifndef MTTS_PSTAGES256
#define MTTS_PSTAGES256 4
#endif
constexpr int kBlockW = 128;      // weight rows per tile == UMMA M
constexpr int kSwizzleBytes = 128;
constexpr int kNumThreads = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2-5: epilogue

struct GemmParams {
  int M, N, K;
  int kb_total;      // ceil(K / BLOCK_K)
  int kb_per_split;  // k-blocks handled by each grid.z slice
  int splits;
  void* out;
  long long ldo;
  int out_bf16;  // 1: bf16 output, 0: fp32 (or fp16, see out_f16)
  int out_f16;   // 1: fp16 output (codec fp16-operand path: no residual / SwiGLU, plain rounding)
  const float* bias;
  const float* gamma;
  const void* residual;
  long long ldr;
  int flags;
  int vec_ok;     // output / residual rows allow 4-wide vector access
  int pair_tiles_n;  // > 0: CTA-pair kernel, number of 256-row weight tiles
  float* partials;   // split-K with the reduction in the consumer: [splits][M][N] fp32
  // fused LM heads + greedy pick (mtts_heads8_sample): instead of logits, every 32-row quarter of the stacked head matrix
  // reports, per batch row, its best and second-best bf16 logit as sortable keys: [M][n_quarters][2] u32
  uint32_t* argmax_keys;
  int n_quarters;
  int n_chan;
  int chan_lo[8], chan_hi[8];  // real rows of channel c in the stacked matrix: [chan_lo, chan_hi)
};

template <typename T>
struct Traits;
template <>
struct Traits<bf16> {
  static constexpr int kBlockK = 64, kUmmaK = 16, kFmt = 1;
};
template <>
struct Traits<float> {
  static constexpr int kBlockK = 32, kUmmaK = 8, kFmt = 2;
};
// fp16 operands (kind::f16, 10-bit mantissa like TF32 at twice the tensor rate and half the operand bytes): the codec's
// large GEMMs, whose inputs are LayerNorm / GELU outputs and weights well inside the fp16 range
template <>
struct Traits<__half> {
  static constexpr int kBlockK = 64, kUmmaK = 16, kFmt = 0;
};

template <int BN>
constexpr uint32_t tmem_cols() {
  return BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
}

template <int BN, int kStages>
constexpr int smem_bytes() {
  return kStages * (kBlockW * kSwizzleBytes + BN * kSwizzleBytes) + 1024 /*align slack*/ + 256 /*barriers*/;
}

// One output row segment of 4 consecutive n for row m: the fused epilogue on a float4 of accumulators.
__device__ __forceinline__ void epilogue_store4(const GemmParams& p, int m, int n, float4 a) {
  float v[4] = {a.x, a.y, a.z, a.w};
  if (p.flags & MTTS_EPI_SWIGLU) {
    // rows interleaved (2j = gate_j, 2j+1 = up_j); rounding points of Qwen3MLP in bf16:
    // bf16(gate), bf16(silu), bf16(up), bf16(product)
    float o2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float g = v[2 * h], u = v[2 * h + 1];
      if (p.out_bf16) { g = bf16_round(g); u = bf16_round(u); }
      float s = silu_nb(g);  // branch-free: the two evaluations of a call (and the unrolled rows around it) interleave
      if (p.out_bf16) s = bf16_round(s);
      o2[h] = s * u;
    }
    const long long o = (long long)m * p.ldo + (n >> 1);
    if (p.out_bf16) {
      bf16* op = reinterpret_cast<bf16*>(p.out) + o;
      if (n + 1 < p.N) op[0] = __float2bfloat16_rn(o2[0]);
      if (n + 3 < p.N) op[1] = __float2bfloat16_rn(o2[1]);
    } else {
      float* op = reinterpret_cast<float*>(p.out) + o;
      if (n + 1 < p.N) op[0] = o2[0];
      if (n + 3 < p.N) op[1] = o2[1];
    }
    return;
  }
  const bool full = p.vec_ok && (n + 3 < p.N);
  float r[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.flags & MTTS_EPI_RESIDUAL) {
    const long long ri = (long long)m * p.ldr + n;
    if (full) {
      if (p.out_bf16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.residual) + ri);
        r[0] = bf16lo(u.x); r[1] = bf16hi(u.x); r[2] = bf16lo(u.y); r[3] = bf16hi(u.y);
      } else {
        const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + ri);
        r[0] = u.x; r[1] = u.y; r[2] = u.z; r[3] = u.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.N)
          r[j] = p.out_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.residual)[ri + j])
                            : reinterpret_cast<const float*>(p.residual)[ri + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n + j >= p.N) break;
    float x = v[j];
    if (p.flags & MTTS_EPI_BIAS) x += __ldg(p.bias + n + j);
    if (p.flags & MTTS_EPI_GELU) x = (p.out_bf16 || (p.flags & MTTS_EPI_EXACT_ACT)) ? gelu_erf(x) : gelu_fast(x);
    if (p.out_bf16) x = bf16_round(x);  // the reference materialises the bf16 linear output first
    if (p.flags & MTTS_EPI_GAMMA) x *= __ldg(p.gamma + n + j);
    if (p.flags & MTTS_EPI_RESIDUAL) x = r[j] + x;
    v[j] = x;
  }
  const long long o = (long long)m * p.ldo + n;
  if (p.out_f16) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < p.N) reinterpret_cast<__half*>(p.out)[o + j] = __float2half_rn(v[j]);
    return;
  }
  if (full) {
    if (p.out_bf16)
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + o) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    else
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < p.N) {
        if (p.out_bf16)
          reinterpret_cast<bf16*>(p.out)[o + j] = __float2bfloat16_rn(v[j]);
        else
          reinterpret_cast<float*>(p.out)[o + j] = v[j];
      }
  }
}

#ifdef MTTS_GEMM_TRACE
// debug build only (scripts/trace_gemm.py): per-CTA phase timestamps (globaltimer, ns) of the most recent launch
__device__ unsigned long long g_gemm_trace[4096 * 16];
__device__ __forceinline__ void gemm_trace(int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (cta < 4096) g_gemm_trace[cta * 16 + slot] = t;
}
#define GEMM_TRACE(slot) gemm_trace(slot)
#else
#define GEMM_TRACE(slot)
#endif

template <typename T, int BN, int kStages>
__global__ void __launch_bounds__(kNumThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                              const __grid_constant__ CUtensorMap tmap_x,
                                                              const GemmParams p) {
  constexpr int BK = Traits<T>::kBlockK;
  constexpr int UK = Traits<T>::kUmmaK;
  constexpr uint32_t kABytes = kBlockW * kSwizzleBytes;
  constexpr uint32_t kBBytes = BN * kSwizzleBytes;
  constexpr uint32_t kCols = tmem_cols<BN>();
  constexpr uint32_t kIdesc = make_idesc(Traits<T>::kFmt, kBlockW, BN);
  static_assert(kStages * (kABytes + kBBytes) >= BN * kBlockW * 4, "ring too small to stage the accumulator tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;  // host guarantees >= 1

  if (threadIdx.x == 0) GEMM_TRACE(0);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kCols>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) GEMM_TRACE(1);
  // Only now (TMEM columns are ours) may the next kernel in the stream start its prologue / weight prefetch: a
  // dependent that grabbed TMEM first and then sat in griddepcontrol.wait would deadlock against our allocation.
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // weights are streamed once per launch when M fits one tile; keep activations resident in L2.
      const uint64_t pol_w = (gridDim.y == 1) ? kEvictFirst : kEvictNormal;
      const uint64_t pol_x = kEvictLast;
      // (1) weights of the first ring fill do not depend on the previous kernel: issue them before the PDL wait
      const int pre = min(num_kb, kStages);
      for (int it = 0; it < pre; ++it) {
        mbar_arrive_expect_tx(&full_bar[it], kABytes + kBBytes);
        tma_load_2d(smem_a + it * kABytes, &tmap_w, &full_bar[it], (kb_begin + it) * BK, n_tile * kBlockW, pol_w);
      }
      pdl_wait();
      GEMM_TRACE(2);
      // (2) activations are the predecessor's output
      for (int it = 0; it < pre; ++it)
        tma_load_2d(smem_b + it * kBBytes, &tmap_x, &full_bar[it], (kb_begin + it) * BK, m_tile * BN, pol_x);
      for (int it = pre; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kABytes + kBBytes);
        const int kc = (kb_begin + it) * BK;
        tma_load_2d(smem_a + s * kABytes, &tmap_w, &full_bar[s], kc, n_tile * kBlockW, pol_w);
        tma_load_2d(smem_b + s * kBBytes, &tmap_x, &full_bar[s], kc, m_tile * BN, pol_x);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        if (it == 0) GEMM_TRACE(3);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * kABytes);
        const uint32_t b_addr = smem_u32(smem_b + s * kBBytes);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * UK * (int)sizeof(T));
          const uint64_t db = make_smem_desc_sw128(b_addr + k * UK * (int)sizeof(T));
          if constexpr (sizeof(T) == 2)
            umma_bf16(tmem_base, da, db, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
          else
            umma_tf32(tmem_base, da, db, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
      GEMM_TRACE(4);
    }
  } else {
    // ================= epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =================
    const int quarter = warp & 3;
    const int n_local = quarter * 32 + lane;
    const int epi_tid = threadIdx.x - 64;  // 0..127
    const int m_base = m_tile * BN;
    const int mv = min(BN, p.M - m_base);  // valid activation rows in this tile
    pdl_wait();  // residual reads and output writes below depend on / conflict with the predecessor
    if (epi_tid == 0) GEMM_TRACE(11);
    mbar_wait(tmem_full_bar, 0);
    if (epi_tid == 0) GEMM_TRACE(5);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // ---- stage the fp32 accumulator (partial) tile as [m][128 n] in the now idle smem ring
    float* stage = reinterpret_cast<float*>(smem);
#pragma unroll 1
    for (int c = 0; c < mv; c += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(taddr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) stage[(c + j) * kBlockW + n_local] = __uint_as_float(r[j]);
    }
    if (epi_tid == 0) GEMM_TRACE(6);
    if (p.splits == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  // ---- split-K: all partial tiles of this cluster are staged
  __syncwarp();  // reconverge the single-lane producer / MMA warps: barrier.cluster is warp-aligned
  if (p.splits > 1) cluster_sync_all();
  if (threadIdx.x == 64) GEMM_TRACE(7);

  if (warp >= 2) {
    const int epi_tid = threadIdx.x - 64;
    const int m_base = m_tile * BN;
    const int mv = min(BN, p.M - m_base);
    const int n4 = (epi_tid & 31) * 4;
    const int n = n_tile * kBlockW + n4;
    const float* stage = reinterpret_cast<const float*>(smem);
    if (n < p.N) {
      if (p.splits > 1) {
        // this CTA finalises rows [r0, r1) of the tile; thread -> 4 consecutive n, rows r0 + (epi_tid / 32) + 4 i
        const int rank = (int)cluster_ctarank();
        const int per = (mv + p.splits - 1) / p.splits;
        const int r0 = rank * per, r1 = min(mv, r0 + per);
        const uint32_t sbase = smem_u32(stage) + n4 * 4;
#pragma unroll 1
        for (int ml = r0 + (epi_tid >> 5); ml < r1; ml += 8) {
          float4 v[2][8];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int s = 0; s < 8; ++s)
              v[i][s] = (s < p.splits && ml + 4 * i < r1) ? ld_dsmem_f4(sbase + (ml + 4 * i) * kBlockW * 4, s)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float4 acc = v[i][0];
#pragma unroll
            for (int s = 1; s < 8; ++s) {  // fixed rank order -> bitwise deterministic
              acc.x += v[i][s].x; acc.y += v[i][s].y; acc.z += v[i][s].z; acc.w += v[i][s].w;
            }
            if (ml + 4 * i < r1) epilogue_store4(p, m_base + ml + 4 * i, n, acc);
          }
        }
      } else {
#pragma unroll 2
        for (int ml = epi_tid >> 5; ml < mv; ml += 4) {
          const float4 a = *reinterpret_cast<const float4*>(stage + ml * kBlockW + n4);
          epilogue_store4(p, m_base + ml, n, a);
        }
      }
    }
  }

  // nobody may exit (and have its shared memory reassigned) while a peer is still reading it. Execution barrier only:
  // the peers' DSMEM loads have returned before they arrive, and our output stores need not be visible to anyone here
  // (pc sampling: the epilogue warps spent ~3 us in the release form of this barrier, waiting for their own stores).
  if (threadIdx.x == 64) GEMM_TRACE(8);
  // (arriving per thread right after its last DSMEM load, before the output stores, was slower: the unaligned barrier
  // forms cost more than the overlap gained)
  __syncwarp();
  if (p.splits > 1) cluster_sync_relaxed();
  if (threadIdx.x == 64) GEMM_TRACE(9);

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
    if (lane == 0) GEMM_TRACE(10);
  }
}

// ---------------------------------------------------------------------------------------------
// Split-K with the reduction moved into the CONSUMER kernel (decode-step projections, M <= 256).
//
// At M = 129..256 a decode projection is no longer bound by HBM but by what an SM can pull over the L2 -> SM crossbar:
// per CTA the weight slice (128 rows x K/S), the activation slice (M rows x K/S) AND the fp32 partial tile (128 x M x 4 B,
// as large as the weight slice) all cross it. The cluster variant above moves the partial tiles twice more (DSMEM reads
// by the reducing CTAs) behind two cluster-wide barriers, which is more than half of every launch.
// Here a CTA owns ALL batch rows of its (weight tile, k-slice) — one 128 x BN accumulator, BN up to 256, so weights
// cross the crossbar exactly once — and simply stores its fp32 partial tile to an L2-resident workspace
// [split][m][n] with coalesced 128-byte warp stores: no cluster, no barrier, no DSMEM. The kernel that consumes the
// projection anyway (residual add + RMSNorm for o_proj / down_proj: mtts_splitk_reduce_rmsnorm; the bf16 cast in front
// of attention for q/k/v: mtts_splitk_reduce) sums the S slices in fixed order (bitwise deterministic) while it reads
// its input, spread over all SMs.
// ---------------------------------------------------------------------------------------------
template <int BN, int kStages>
__global__ void __launch_bounds__(kNumThreads) gemm_tc_partial_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                      const __grid_constant__ CUtensorMap tmap_x,
                                                                      const GemmParams p) {
  using T = bf16;
  constexpr int BK = Traits<T>::kBlockK;
  constexpr int UK = Traits<T>::kUmmaK;
  constexpr uint32_t kABytes = kBlockW * kSwizzleBytes;
  constexpr uint32_t kBBytes = BN * kSwizzleBytes;
  constexpr uint32_t kCols = tmem_cols<BN>();
  constexpr uint32_t kIdesc = make_idesc(Traits<T>::kFmt, kBlockW, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;  // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kCols>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();  // after the TMEM allocation (see gemm_tc_kernel)

  if (warp == 0) {
    if (lane == 0) {
      const int pre = min(num_kb, kStages);
      for (int it = 0; it < pre; ++it) {  // weights never depend on the predecessor: stream them before the PDL wait
        mbar_arrive_expect_tx(&full_bar[it], kABytes + kBBytes);
        tma_load_2d(smem_a + it * kABytes, &tmap_w, &full_bar[it], (kb_begin + it) * BK, n_tile * kBlockW, kEvictFirst);
      }
      pdl_wait();
      for (int it = 0; it < pre; ++it)
        tma_load_2d(smem_b + it * kBBytes, &tmap_x, &full_bar[it], (kb_begin + it) * BK, m_tile * BN, kEvictLast);
      for (int it = pre; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kABytes + kBBytes);
        const int kc = (kb_begin + it) * BK;
        tma_load_2d(smem_a + s * kABytes, &tmap_w, &full_bar[s], kc, n_tile * kBlockW, kEvictFirst);
        tma_load_2d(smem_b + s * kBBytes, &tmap_x, &full_bar[s], kc, m_tile * BN, kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * kABytes);
        const uint32_t b_addr = smem_u32(smem_b + s * kBBytes);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * UK * (int)sizeof(T));
          const uint64_t db = make_smem_desc_sw128(b_addr + k * UK * (int)sizeof(T));
          umma_bf16(tmem_base, da, db, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- epilogue warps 2..5 (TMEM lane quarters 2,3,0,1): lane = weight row n, TMEM column = batch row m
    const int quarter = warp & 3;
    const int n = n_tile * kBlockW + quarter * 32 + lane;
    const int m_base = m_tile * BN;
    const int mv = min(BN, p.M - m_base);
    pdl_wait();  // the workspace may still be read by the predecessor's predecessor's consumer: write only after it
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* dst = p.partials + ((long long)split * p.M + m_base) * p.N + n;
    const bool n_ok = n < p.N;
#pragma unroll 1
    for (int c = 0; c < mv; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c, r);
      tmem_ld_wait();
      if (n_ok) {
        if (c + 32 <= mv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(long long)(c + j) * p.N] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c + j < mv) dst[(long long)(c + j) * p.N] = __uint_as_float(r[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for large M (prefill, codec): one CTA per SM walks the (n_tile, m_tile) list; the 512 TMEM
// columns hold TWO 128 x 256 fp32 accumulators, so the MMA warp fills one while eight epilogue warps drain the
// other straight from TMEM to global memory (lane = output column n: every warp store is one contiguous segment of
// row m). The smem ring keeps streaming across tile boundaries.
// ---------------------------------------------------------------------------------------------
constexpr int kPBN = 256;
constexpr int kPStages = 4;
constexpr int kPThreads = 320;  // warp0 TMA, warp1 MMA/TMEM, warps 2..9 epilogue

__device__ __forceinline__ float epilogue_scalar(const GemmParams& p, float v, int m, int n, float bias_n, float gamma_n) {
  if (p.flags & MTTS_EPI_BIAS) v += bias_n;
  if (p.flags & MTTS_EPI_GELU) v = (p.out_bf16 || (p.flags & MTTS_EPI_EXACT_ACT)) ? gelu_erf(v) : gelu_fast(v);
  if (p.out_bf16) v = bf16_round(v);
  if (p.flags & MTTS_EPI_GAMMA) v *= gamma_n;
  if (p.flags & MTTS_EPI_RESIDUAL) {
    const long long ri = (long long)m * p.ldr + n;
    const float r = p.out_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.residual)[ri])
                               : reinterpret_cast<const float*>(p.residual)[ri];
    v = r + v;
  }
  return v;
}

// SwiGLU epilogue of one TMEM chunk (32 batch rows x this lane's weight row; rows interleaved: even lane = gate row of
// output column n / 2, odd lane = its up row). Lane pairs exchange ONE value per two batch rows, so that the even lane
// finishes batch row j and the odd lane batch row j + 1: all 32 lanes work. Written as four straight-line phases (round,
// exchange, activate, store) over register arrays: the warp issues in order, so 16 independent ~200-cycle chains
// (cvt -> shfl -> ex2 -> rcp -> fma -> cvt) only overlap if the code interleaves them. The first form — per row: shuffle,
// then expf and an IEEE division behind a branch on the even lanes — ran them back to back: 23 us to drain one
// 128 x 256 tile per SM, more than the tile's MMAs take (globaltimer phase trace of the drain, round 2).
template <bool kBf16>
__device__ __forceinline__ void swiglu_chunk(const GemmParams& p, const uint32_t (&r)[32], int mrow0, int n, bool n_ok, int lane) {
  const bool odd = lane & 1;
  float keep[16], send[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = __uint_as_float(r[2 * i]), b = __uint_as_float(r[2 * i + 1]);
    if (kBf16) {  // bf16(gate), bf16(up): one packed conversion for the two rows
      const __nv_bfloat162 ab = __floats2bfloat162_rn(a, b);
      a = __low2float(ab);
      b = __high2float(ab);
    }
    keep[i] = odd ? b : a;  // even lane: gate of row 2i; odd lane: up of row 2i + 1
    send[i] = odd ? a : b;  // even lane: gate of row 2i + 1 (for the odd lane); odd lane: up of row 2i (for the even lane)
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) send[i] = __shfl_xor_sync(0xffffffffu, send[i], 1);
  float h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float g = odd ? send[i] : keep[i], u = odd ? keep[i] : send[i];
    float sg = silu_nb(g);
    if (kBf16) sg = bf16_round(sg);
    h[i] = sg * u;
  }
  const int m = mrow0 + (odd ? 1 : 0);
  const long long o = (long long)m * p.ldo + (n >> 1);
  if (n_ok) {
    if (kBf16) {
      bf16* op = reinterpret_cast<bf16*>(p.out) + o;
      if (mrow0 + 32 <= p.M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) op[(long long)(2 * i) * p.ldo] = __float2bfloat16_rn(h[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (m + 2 * i < p.M) op[(long long)(2 * i) * p.ldo] = __float2bfloat16_rn(h[i]);
      }
    } else {
      float* op = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (m + 2 * i < p.M) op[(long long)(2 * i) * p.ldo] = h[i];
    }
  }
}

// Drains 32 TMEM lanes x 128 accumulator columns (one epilogue warp's share of a tile) through the fused epilogue:
// lane = output column n, TMEM column = activation row m0 + c.
// `release_bar` != 0 (CTA-pair kernel): cluster address of the leader's tmem_empty barrier; the warp arrives on it as
// soon as its LAST tcgen05.ld has returned — before the arithmetic and the global stores of that chunk — so the MMA warp
// gets the accumulator back while the epilogue is still writing.
__device__ __forceinline__ void persist_drain(const GemmParams& p, uint32_t taddr, int m0, int n, bool n_ok, float bias_n,
                                              float gamma_n, int lane, uint32_t release_bar = 0) {
  if (p.argmax_keys) {
    // lane = vocabulary row n (one 32-row quarter per warp, never straddling two channels: heads are padded to 32 rows),
    // TMEM column = batch row. key = (order-preserving 16-bit image of the bf16 logit) << 16 | (31 - lane): the warp-wide
    // integer maximum is the best logit, the lowest row on ties (torch.argmax); 0 marks padding rows / "no candidate".
    bool real = false;
    if (n_ok)
      for (int ch = 0; ch < p.n_chan; ++ch) real = real || (n >= p.chan_lo[ch] && n < p.chan_hi[ch]);
    const int q = n >> 5;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      if (m0 + c >= p.M) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c, r);
      tmem_ld_wait();
      uint32_t best_mine = 0u, second_mine = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t hb = __float_as_uint(bf16_round(__uint_as_float(r[j]))) >> 16;  // full-rate F2FP, not F2F
        const uint32_t ord = (hb & 0x8000u) ? (~hb & 0xffffu) : (hb | 0x8000u);
        const uint32_t key = real ? ((ord << 16) | (uint32_t)(31 - lane)) : 0u;
        const uint32_t best = __reduce_max_sync(0xffffffffu, key);
        const uint32_t second = __reduce_max_sync(0xffffffffu, key == best ? 0u : key);
        if (lane == j) { best_mine = best; second_mine = second; }
      }
      const int m = m0 + c + lane;
      if (m < p.M && q < p.n_quarters)  // the last tile's quarters beyond N must not alias the next row's first slots
        *reinterpret_cast<uint2*>(p.argmax_keys + ((long long)m * p.n_quarters + q) * 2) = make_uint2(best_mine, second_mine);
    }
    if (release_bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(release_bar);
    }
    return;
  }
  bool released = false;
#pragma unroll 1
  for (int c = 0; c < 128; c += 32) {
    if (m0 + c >= p.M) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + c, r);
    tmem_ld_wait();
    if (release_bar && (c == 96 || m0 + c + 32 >= p.M)) {  // last chunk of this warp: the accumulator is free
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(release_bar);
      released = true;
    }
    if (p.flags & MTTS_EPI_SWIGLU) {
      if (p.out_bf16)
        swiglu_chunk<true>(p, r, m0 + c, n, n_ok, lane);
      else
        swiglu_chunk<false>(p, r, m0 + c, n, n_ok, lane);
    } else if (p.out_f16 && (p.flags & ~(MTTS_EPI_BIAS | MTTS_EPI_GELU)) == 0) {
      // fp16-operand codec path: the GELU'd intermediate goes straight out as fp16 (half the bytes of the largest tensor)
      // lane pairs exchange one value per two rows so that every lane stores a half2 (even lanes row j, odd lanes row
      // j + 1): 16 four-byte store instructions per 32 rows instead of 32 two-byte ones (1660 -> ~1130 us on the
      // ConvNeXt pw1 shape, where the epilogue, not the MMA, sets the pace)
      const int lim = min(32, p.M - (m0 + c));
      const int odd = lane & 1;
      if ((p.ldo & 1) == 0 && (p.N & 1) == 0) {
        // straight-line phases (convert, activate, exchange, store). The arithmetic runs on PACKED fp16 pairs — rows
        // 2i, 2i + 1 of this lane's column share a register — so the GELU costs 9 instructions per pair (gelu_tanh5_h2),
        // and one shuffle + one byte permute per pair turns (two rows of one column) into (two columns of one row).
        __half* op = reinterpret_cast<__half*>(p.out) + (long long)(m0 + c + odd) * p.ldo + (n - odd);
        __half2 h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = __floats2half2_rn(__uint_as_float(r[2 * i]) + bias_n, __uint_as_float(r[2 * i + 1]) + bias_n);
        if (p.flags & MTTS_EPI_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) h[i] = gelu_tanh5_h2(h[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t mine = *reinterpret_cast<const uint32_t*>(&h[i]);
          const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
          // even lane: row 2i of columns (n, n + 1) = (mine.lo, other.lo); odd lane: row 2i + 1 of (n - 1, n) = (other.hi, mine.hi)
          const uint32_t o = odd ? __byte_perm(mine, other, 0x3276) : __byte_perm(mine, other, 0x5410);
          h[i] = *reinterpret_cast<const __half2*>(&o);
        }
        if (n_ok) {
          if (lim == 32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) *reinterpret_cast<__half2*>(op + (long long)(2 * i) * p.ldo) = h[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (2 * i + odd < lim) *reinterpret_cast<__half2*>(op + (long long)(2 * i) * p.ldo) = h[i];
          }
        }
      } else if (n_ok) {
        __half* op = reinterpret_cast<__half*>(p.out) + (long long)(m0 + c) * p.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < lim) {
            const float x = __uint_as_float(r[j]) + bias_n;
            op[(long long)j * p.ldo] = __float2half_rn((p.flags & MTTS_EPI_GELU) ? gelu_tanh5(x) : x);
          }
      }
    } else if (p.flags == (MTTS_EPI_BIAS | MTTS_EPI_GELU) && !p.out_bf16) {
      // the codec's MLP up-projections: the flag tests are hoisted, nothing but bias + GELU + one store per element
      float* op = reinterpret_cast<float*>(p.out) + (long long)(m0 + c) * p.ldo + n;
      if (n_ok) {
        if (m0 + c + 32 <= p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) op[(long long)j * p.ldo] = gelu_fast(__uint_as_float(r[j]) + bias_n);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (m0 + c + j < p.M) op[(long long)j * p.ldo] = gelu_fast(__uint_as_float(r[j]) + bias_n);
        }
      }
    } else if (!p.out_bf16 && !p.out_f16 && !(p.flags & MTTS_EPI_GELU)) {
      // fp32 bias / layer-scale / residual (the codec's down-projections and attention outputs). The residual may
      // alias the output (x += ...), so its 32 loads are issued explicitly BEFORE the first store: left to the
      // compiler every load waits behind the previous element's store (one L2 round trip per element).
      if (n_ok) {
        float* op = reinterpret_cast<float*>(p.out) + (long long)(m0 + c) * p.ldo + n;
        const int lim = min(32, p.M - (m0 + c));
        float res[32];
        if (p.flags & MTTS_EPI_RESIDUAL) {
          const float* rp = reinterpret_cast<const float*>(p.residual) + (long long)(m0 + c) * p.ldr + n;
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = j < lim ? __ldcg(rp + (long long)j * p.ldr) : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < lim) op[(long long)j * p.ldo] = fmaf(__uint_as_float(r[j]) + bias_n, gamma_n, res[j]);
      }
    } else if (p.out_bf16 && !(p.flags & MTTS_EPI_GELU)) {
      // bf16 linear (+ residual): prefill projections. Lane pairs exchange one value per two rows so that every lane
      // stores (and, for the residual, loads) a bf16x2: even lanes finish row j, odd lanes row j + 1 of columns
      // (n & ~1, n | 1) — 16 four-byte accesses per 32 rows instead of 32 two-byte ones, in straight-line phases. The
      // residual may alias the output (x += ...): all of its loads are issued before the first store.
      const int lim = min(32, p.M - (m0 + c));
      const bool has_res = (p.flags & MTTS_EPI_RESIDUAL) != 0;
      const int odd = lane & 1;
      const bool pairs_ok = (p.ldo & 1) == 0 && (p.N & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 3) == 0 &&
                            (!has_res || ((p.ldr & 1) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 3) == 0));
      if (pairs_ok) {
        float g[32], rv[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = bf16_round(__uint_as_float(r[j]) + bias_n) * gamma_n;  // the bf16 linear output first
#pragma unroll
        for (int i = 0; i < 16; ++i) rv[i] = __shfl_xor_sync(0xffffffffu, odd ? g[2 * i] : g[2 * i + 1], 1);
        uint32_t res[16];
        if (has_res && n_ok) {
          const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + (long long)(m0 + c + odd) * p.ldr + (n - odd);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            res[i] = (2 * i + odd < lim) ? __ldcg(reinterpret_cast<const unsigned int*>(rp + (long long)(2 * i) * p.ldr)) : 0u;
        }
        if (n_ok) {
          bf16* op = reinterpret_cast<bf16*>(p.out) + (long long)(m0 + c + odd) * p.ldo + (n - odd);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float lo = odd ? rv[i] : g[2 * i], hi = odd ? g[2 * i + 1] : rv[i];
            if (has_res) {
              lo += bf16lo(res[i]);
              hi += bf16hi(res[i]);
            }
            if (2 * i + odd < lim) *reinterpret_cast<uint32_t*>(op + (long long)(2 * i) * p.ldo) = pack_bf16(lo, hi);
          }
        }
      } else if (n_ok) {
        bf16* op = reinterpret_cast<bf16*>(p.out) + (long long)(m0 + c) * p.ldo + n;
        float res[32];
        if (has_res) {
          const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + (long long)(m0 + c) * p.ldr + n;
#pragma unroll
          for (int j = 0; j < 32; ++j) res[j] = j < lim ? __bfloat162float(__ldcg(rp + (long long)j * p.ldr)) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < lim) {
            float v = bf16_round(__uint_as_float(r[j]) + bias_n) * gamma_n;  // the bf16 linear output comes first
            if (has_res) v += res[j];
            op[(long long)j * p.ldo] = __float2bfloat16_rn(v);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + c + j;
        if (n_ok && m < p.M) {
          const float v = epilogue_scalar(p, __uint_as_float(r[j]), m, n, bias_n, gamma_n);
          const long long o = (long long)m * p.ldo + n;
          if (p.out_f16)
            reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(v);
          else if (p.out_bf16)
            reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16_rn(v);
          else
            reinterpret_cast<float*>(p.out)[o] = v;
        }
      }
    }
  }
  if (release_bar && !released) {  // nothing to drain (rows beyond M)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster_relaxed(release_bar);
  }
}

template <typename T>
__global__ void __launch_bounds__(kPThreads, 1) gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                       const __grid_constant__ CUtensorMap tmap_x,
                                                                       const GemmParams p, int tiles_n, int num_tiles) {
  constexpr int BK = Traits<T>::kBlockK;
  constexpr int UK = Traits<T>::kUmmaK;
  constexpr uint32_t kABytes = kBlockW * kSwizzleBytes;
  constexpr uint32_t kBBytes = kPBN * kSwizzleBytes;
  constexpr uint32_t kIdesc = make_idesc(Traits<T>::kFmt, kBlockW, kPBN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kPStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kPStages * kBBytes);
  uint64_t* empty_bar = full_bar + kPStages;
  uint64_t* tmem_full_bar = empty_bar + kPStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.kb_total;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kPStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full_bar[a], 1);
        mbar_init(&tmem_empty_bar[a], 8);  // one arrival per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();  // after the TMEM allocation (see gemm_tc_kernel)

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      pdl_wait();
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % tiles_n, m_tile = t / tiles_n;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], kABytes + kBBytes);
          tma_load_2d(smem_a + s * kABytes, &tmap_w, &full_bar[s], kb * BK, n_tile * kBlockW, kEvictLast);
          tma_load_2d(smem_b + s * kBBytes, &tmap_x, &full_bar[s], kb * BK, m_tile * kPBN, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1;
        mbar_wait(&tmem_empty_bar[a], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + a * kPBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kPStages;
          const uint32_t ph = (it / kPStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + s * kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UK * (int)sizeof(T));
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UK * (int)sizeof(T));
            if constexpr (sizeof(T) == 2)
              umma_bf16(tacc, da, db, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
            else
              umma_tf32(tacc, da, db, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[a]);
      }
    }
  } else {
    // ================= epilogue: 8 warps, lane quarter = warp % 4, column half = (warp - 2) / 4 =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    pdl_wait();
    int ti = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
      const int n_tile = t % tiles_n, m_tile = t / tiles_n;
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;
      const int n = n_tile * kBlockW + quarter * 32 + lane;
      const bool n_ok = n < p.N;
      const float bias_n = (n_ok && (p.flags & MTTS_EPI_BIAS)) ? __ldg(p.bias + n) : 0.f;
      const float gamma_n = (n_ok && (p.flags & MTTS_EPI_GAMMA)) ? __ldg(p.gamma + n) : 1.f;
      mbar_wait(&tmem_full_bar[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * kPBN + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m0 = m_tile * kPBN + half * 128;
      persist_drain(p, taddr, m0, n, n_ok, bias_n, gamma_n, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[a]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel (cta_group::2): the two CTAs of a 2-cluster (one TPC) share a
// 256 (weight rows) x 256 (activation rows) tile. Each CTA stages only ITS 128 weight rows and ITS 128 activation rows
// per k-block — 32 KB instead of the 48 KB a single-CTA 128 x 256 tile needs for the same 2 x 128 x 256 x BK flops per
// SM — and the leader's MMA reads both halves of the activation tile, one from each SM. At TF32 the single-CTA kernel
// is bound by what one SM can pull from L2 (48 KB per 2.1 MFLOP, ~520 TFLOP/s at the ~80 GB/s per SM the crossbar
// sustains); the pair needs a third less. Accumulators: 128 lanes x 256 columns per CTA, two of them (double buffer).
// ---------------------------------------------------------------------------------------------
constexpr int kP2Stages = 6;  // 6 x 32 KB (5 and 7 measure the same within run-to-run noise)
constexpr int kP2Bytes = kP2Stages * 2 * kBlockW * kSwizzleBytes + 1024 + 256;

template <typename T>
__global__ void __launch_bounds__(kPThreads, 1) gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                    const __grid_constant__ CUtensorMap tmap_x,
                                                                    const GemmParams p, int tiles_n, int num_tiles) {
  constexpr int BK = Traits<T>::kBlockK;
  constexpr int UK = Traits<T>::kUmmaK;
  constexpr uint32_t kHalfBytes = kBlockW * kSwizzleBytes;  // 128 rows x 128 B: one operand half per CTA per stage
  constexpr uint32_t kIdesc = make_idesc(Traits<T>::kFmt, 2 * kBlockW, kPBN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kP2Stages * kHalfBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kP2Stages * kHalfBytes);  // used in the leader only
  uint64_t* empty_bar = full_bar + kP2Stages;
  uint64_t* tmem_full_bar = empty_bar + kP2Stages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2], used in the leader only: 16 epilogue warps arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.kb_total;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kP2Stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full_bar[a], 1);
        mbar_init(&tmem_empty_bar[a], 16);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm<512>(tmem_ptr_smem);
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before either one signals the other
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own halves, byte counts land on the leader's full barrier =================
    if (lane == 0) {
      const uint32_t leader_full = mapa_u32(smem_u32(full_bar), 0);
      // weights never depend on the predecessor kernel: the first ring fill of the first tile is requested before the
      // PDL wait (a decode-step gate/up projection then starts streaming while the RMSNorm in front of it still runs)
      const uint64_t pol_w = num_tiles <= num_pairs ? kEvictFirst : kEvictLast;  // one tile per pair: weights are read once
      int pre = 0;
      if (pair < num_tiles) {
        pre = min(num_kb, kP2Stages);
        const int n_tile0 = pair % tiles_n;
        for (int kb = 0; kb < pre; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[kb], 4 * kHalfBytes);
          tma_load_2d_2sm(smem_a + kb * kHalfBytes, &tmap_w, leader_full + kb * 8, kb * BK,
                          n_tile0 * 2 * kBlockW + (int)rank * kBlockW, pol_w);
        }
      }
      pdl_wait();
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int n_tile = t % tiles_n, m_tile = t / tiles_n;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kP2Stages;
          const uint32_t ph = (it / kP2Stages) & 1;
          if (it >= pre) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 4 * kHalfBytes);
            tma_load_2d_2sm(smem_a + s * kHalfBytes, &tmap_w, leader_full + s * 8, kb * BK,
                            n_tile * 2 * kBlockW + (int)rank * kBlockW, pol_w);
          }
          tma_load_2d_2sm(smem_b + s * kHalfBytes, &tmap_x, leader_full + s * 8, kb * BK,
                          m_tile * kPBN + (int)rank * (kPBN / 2), kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (lane == 0 && rank == 0) {
      int it = 0, ti = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++ti) {
        const int a = ti & 1;
        const uint32_t aph = (ti >> 1) & 1;
        mbar_wait(&tmem_empty_bar[a], aph ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + a * kPBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kP2Stages;
          const uint32_t ph = (it / kP2Stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * kHalfBytes);
          const uint32_t b_addr = smem_u32(smem_b + s * kHalfBytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UK * (int)sizeof(T));
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UK * (int)sizeof(T));
            if constexpr (sizeof(T) == 2)
              umma_bf16_2sm(tacc, da, db, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
            else
              umma_tf32_2sm(tacc, da, db, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[s], 3);  // frees the stage in both CTAs
        }
        umma_commit_2sm(&tmem_full_bar[a], 3);
      }
    }
  } else {
    // ================= epilogue (both CTAs): own 128 weight rows x all 256 activation rows =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    pdl_wait();
    const uint32_t leader_empty = mapa_u32(smem_u32(tmem_empty_bar), 0);
    int ti = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++ti) {
      const int n_tile = t % tiles_n, m_tile = t / tiles_n;
      const int a = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;
      const int n = n_tile * 2 * kBlockW + (int)rank * kBlockW + quarter * 32 + lane;
      const bool n_ok = n < p.N;
      const float bias_n = (n_ok && (p.flags & MTTS_EPI_BIAS)) ? __ldg(p.bias + n) : 0.f;
      const float gamma_n = (n_ok && (p.flags & MTTS_EPI_GAMMA)) ? __ldg(p.gamma + n) : 1.f;
      mbar_wait(&tmem_full_bar[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * kPBN + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m0 = m_tile * kPBN + half * 128;
      persist_drain(p, taddr, m0, n, n_ok, bias_n, gamma_n, lane, leader_empty + a * 8);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave (or free its TMEM) while the pair's MMAs or remote arrivals are in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows, elem_bytes;  // elem_bytes: 2 = bf16, 4 = fp32, -2 = fp16
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           elem_bytes == o.elem_bytes;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.rows);
    mix((size_t)k.cols);
    mix((size_t)k.ld);
    mix((size_t)k.box_rows);
    mix((size_t)k.elem_bytes);
    return h;
  }
};

std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int get_tmap(const void* ptr, long long rows, long long cols, long long ld, int box_rows, int elem_code,
             CUtensorMap* out) {
  TmapKey key{ptr, rows, cols, ld, box_rows, elem_code};
  const int elem_bytes = elem_code < 0 ? -elem_code : elem_code;
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return MTTS_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return mtts_set_error(MTTS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(kSwizzleBytes / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, elem_code == -2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                       : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mtts_set_error(MTTS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%d", (int)r,
                          rows, cols, ld, box_rows);
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
    g_tmap_cache[key] = m;
  }
  *out = m;
  return MTTS_OK;
}

// Stage counts: the small-BN (decode) variants keep the ring <= ~100 KB so that this GEMM's CTA and the next
// kernel's (PDL) fit on one SM together; the large-BN (prefill / codec) variants use the full 200 KB.
template <int BN> struct Stages;
template <> struct Stages<16> { static constexpr int v = 5; };   //  90 KB
template <> struct Stages<32> { static constexpr int v = 5; };   // 100 KB
template <> struct Stages<64> { static constexpr int v = 4; };   //  96 KB
template <> struct Stages<128> { static constexpr int v = 3; };  //  96 KB
template <> struct Stages<256> { static constexpr int v = 4; };  // 192 KB

template <typename T, int BN>
int launch(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int kStages = Stages<BN>::v;
  constexpr int smem = smem_bytes<BN, kStages>();
  MTTS_CUDA_CHECK(mtts_launch_cluster(gemm_tc_kernel<T, BN, kStages>, grid, dim3(kNumThreads), smem, stream, (int)grid.z,
                                      tw, tx, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

// Ring depth of the partial-tile kernel: BN <= 128 keeps two CTAs per SM (<= ~100 KB each), BN = 256 owns the SM.
template <int BN> struct PStages;
template <> struct PStages<16> { static constexpr int v = 5; };
template <> struct PStages<32> { static constexpr int v = 5; };
template <> struct PStages<64> { static constexpr int v = 4; };
template <> struct PStages<128> { static constexpr int v = 3; };
template <> struct PStages<256> { static constexpr int v = MTTS_PSTAGES256; };  // x 48 KB

template <int BN>
int launch_partial(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int kStages = PStages<BN>::v;
  MTTS_CUDA_CHECK(mtts_launch(gemm_tc_partial_kernel<BN, kStages>, grid, dim3(kNumThreads), smem_bytes<BN, kStages>(), stream,
                              tw, tx, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

template <int BN>
int configure_partial() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_partial_kernel<BN, PStages<BN>::v>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<BN, PStages<BN>::v>()));
  return MTTS_OK;
}

static bool persist_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTTS_GEMM_NO_PERSIST");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <typename T>
int dispatch(int bn, const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, dim3 grid,
             cudaStream_t stream) {
  switch (bn) {
    case 16: return launch<T, 16>(tw, tx, p, grid, stream);
    case 32: return launch<T, 32>(tw, tx, p, grid, stream);
    case 64: return launch<T, 64>(tw, tx, p, grid, stream);
    case 128: return launch<T, 128>(tw, tx, p, grid, stream);
    default: {
      // a single row of tiles (decode at batch 129..256) is a weight-streaming problem: cluster split-K variant
      if (persist_disabled()) return launch<T, 256>(tw, tx, p, grid, stream);
      if (p.pair_tiles_n > 0) {  // CTA-pair kernel: 256 x 256 tiles, tensor maps with 128-row boxes
        const int num_tiles = p.pair_tiles_n * (int)grid.y;
        const int pairs = num_tiles < mtts_num_sms() / 2 ? num_tiles : mtts_num_sms() / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(kPThreads);
        cfg.dynamicSmemBytes = kP2Bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (mtts_pdl_enabled()) {
          attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[na].val.programmaticStreamSerializationAllowed = 1;
          ++na;
        }
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
        cfg.attrs = attr;
        cfg.numAttrs = na;
        MTTS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<T>, tw, tx, p, p.pair_tiles_n, num_tiles));
        MTTS_LAUNCH_CHECK();
        return MTTS_OK;
      }
      const int num_tiles = (int)(grid.x * grid.y);
      const int ctas = num_tiles < mtts_num_sms() ? num_tiles : mtts_num_sms();
      MTTS_CUDA_CHECK(mtts_launch(gemm_tc_persist_kernel<T>, dim3(ctas), dim3(kPThreads), smem_bytes<kPBN, kPStages>(), stream,
                                  tw, tx, p, (int)grid.x, num_tiles));
      MTTS_LAUNCH_CHECK();
      return MTTS_OK;
    }
  }
}

template <typename T, int BN>
int configure_one() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<T, BN, Stages<BN>::v>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<BN, Stages<BN>::v>()));
  return MTTS_OK;
}

}  // namespace

// 2-D tiled tensor map with 128-byte-swizzled boxes of `box_rows` rows x 128 bytes (cached per pointer / shape); used by
// the other tcgen05 kernels of the library (mha_tc5.cu). elem_code: 2 = bf16, 4 = fp32, -2 = fp16.
int mtts_get_tmap_2d(const void* ptr, long long rows, long long cols, long long ld, int box_rows, int elem_code, CUtensorMap* out) {
  return get_tmap(ptr, rows, cols, ld, box_rows, elem_code, out);
}

// Opt every instantiation into its dynamic shared-memory size (called from mtts_init, outside any graph capture).
int mtts_configure_gemm_tc() {
  int rc = 0;
  if ((rc = configure_one<bf16, 16>())) return rc;
  if ((rc = configure_one<bf16, 32>())) return rc;
  if ((rc = configure_one<bf16, 64>())) return rc;
  if ((rc = configure_one<bf16, 128>())) return rc;
  if ((rc = configure_one<bf16, 256>())) return rc;
  if ((rc = configure_one<float, 16>())) return rc;
  if ((rc = configure_one<float, 32>())) return rc;
  if ((rc = configure_one<float, 64>())) return rc;
  if ((rc = configure_one<float, 128>())) return rc;
  if ((rc = configure_one<float, 256>())) return rc;
  if ((rc = configure_one<__half, 16>())) return rc;
  if ((rc = configure_one<__half, 32>())) return rc;
  if ((rc = configure_one<__half, 64>())) return rc;
  if ((rc = configure_one<__half, 128>())) return rc;
  if ((rc = configure_one<__half, 256>())) return rc;
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persist_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<kPBN, kPStages>()));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_pair_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Bytes));
  if ((rc = configure_partial<16>())) return rc;
  if ((rc = configure_partial<32>())) return rc;
  if ((rc = configure_partial<64>())) return rc;
  if ((rc = configure_partial<128>())) return rc;
  if ((rc = configure_partial<256>())) return rc;
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persist_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<kPBN, kPStages>()));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_persist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<kPBN, kPStages>()));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_pair_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Bytes));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_pair_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Bytes));
  return MTTS_OK;
}

int mtts_gemm_tc_pick_bn(int M) {
  static int force = -1;
  if (force < 0) {
    const char* e = getenv("MTTS_GEMM_BN");
    force = e ? atoi(e) : 0;
  }
  if (force) return force;
  if (M <= 16) return 16;
  if (M <= 32) return 32;
  if (M <= 64) return 64;
  if (M <= 256) return 128;  // decode at batch 65..256: one or two 128-row activation tiles, cluster split-K
  return 256;                // prefill / codec: persistent kernel
}

// Split-K heuristic for the weight-streaming (small-M) variants: cluster sizes 1/2/4/8 (odd cluster sizes place
// badly: s = 3 on the 96-tile gate/up projection measured 22 us against 15.6 us for s = 2), aiming at two CTAs per SM
// (two ring buffers fit), each split keeping at least 2 k-blocks.
static int pick_splits(int tiles, int kb_total, int bn, int tiles_m) {
  if (bn > 128) return 1;  // many-tile problems (prefill, codec) go to the persistent kernel
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MTTS_GEMM_SPLITS");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced >= 8 ? 8 : forced >= 4 ? 4 : forced >= 2 ? 2 : 1;  // cluster sizes are powers of two
  // (A cta_group::2 version of this kernel for batch 129..256 — pairs over two adjacent weight tiles, half the L2->SM
  // bytes — was bit-exact but slower in the full step, 5.05-5.9 ms against 4.90 ms: its extra cluster barriers and the
  // two-pass reduction of the 256 x 128 partial tile cost more than the main loop saved. Moving the partial tiles
  // through an L2 workspace instead of DSMEM, 16 DSMEM loads per round and a non-inlined epilogue were each slower too.)
  // (batch 129..256, two activation tiles per weight tile: raising the target to 3, 4 or 6 CTAs per SM, i.e. splitting
  // the 192-tile gate/up projection 2 or 4 ways and q/k/v 8 ways, loses in the full step: 5.27 / 5.62 / 5.60 ms against
  // 4.97 ms at batch 256.)
  const int target = 2 * mtts_num_sms();
  int s = 1;
  while (s < 8 && tiles * (s * 2) <= target && kb_total / (s * 2) >= 2) s *= 2;
  return s;
}

// Kept for ABI stability: split-K partials now live in distributed shared memory, so no workspace is needed.
extern "C" size_t mtts_gemm_workspace_bytes(int M, int N, int K, int dtype) {
  (void)M; (void)N; (void)K; (void)dtype;
  return 256;
}

extern "C" int mtts_gemm(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                         int M, int N, int K, int in_dtype, int out_dtype, int flags, const float* bias,
                         const float* gamma, const void* residual, long long ldr, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(M > 0 && N > 0 && K > 0, "mtts_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MTTS_REQUIRE(in_dtype == MTTS_DTYPE_BF16 || in_dtype == MTTS_DTYPE_F32 || in_dtype == MTTS_DTYPE_F16,
               "mtts_gemm: bad in_dtype %d", in_dtype);
  MTTS_REQUIRE(out_dtype == MTTS_DTYPE_BF16 || out_dtype == MTTS_DTYPE_F32 || out_dtype == MTTS_DTYPE_F16,
               "mtts_gemm: bad out_dtype %d", out_dtype);
  if (out_dtype == MTTS_DTYPE_F16)
    MTTS_REQUIRE(!(flags & (MTTS_EPI_RESIDUAL | MTTS_EPI_SWIGLU)), "mtts_gemm: fp16 output supports bias / GELU / layer-scale only");
  const int eb = in_dtype == MTTS_DTYPE_F32 ? 4 : 2;
  const int ecode = in_dtype == MTTS_DTYPE_F16 ? -2 : eb;
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
               "mtts_gemm: x and w must be 16-byte aligned");
  MTTS_REQUIRE((ldx * eb) % 16 == 0 && (ldw * eb) % 16 == 0, "mtts_gemm: row strides must be multiples of 16 bytes");
  MTTS_REQUIRE(ldx >= K && ldw >= K, "mtts_gemm: leading dimensions smaller than K");
  if (flags & MTTS_EPI_BIAS) MTTS_REQUIRE(bias != nullptr, "mtts_gemm: EPI_BIAS without bias");
  if (flags & MTTS_EPI_GAMMA) MTTS_REQUIRE(gamma != nullptr, "mtts_gemm: EPI_GAMMA without gamma");
  if (flags & MTTS_EPI_RESIDUAL) MTTS_REQUIRE(residual != nullptr, "mtts_gemm: EPI_RESIDUAL without residual");
  if (flags & MTTS_EPI_SWIGLU)
    MTTS_REQUIRE((N % 4) == 0 && !(flags & ~(MTTS_EPI_SWIGLU | MTTS_EPI_EXACT_ACT)),
                 "mtts_gemm: SWIGLU needs N % 4 == 0 and no other flags");

  int bn = mtts_gemm_tc_pick_bn(M);
  // (Routing wide matrices at batch 129..256 to the persistent 256-row-tile kernel wins in isolation — 21 vs 30 us at
  // N = 12288 — but loses inside a decode step, 5.31 vs 5.23 ms at batch 256: its one-CTA-per-SM footprint keeps the
  // neighbouring launches from becoming resident early and prefetching their weights.)
  // The LM heads at batch 65..256 (1250 weight tiles) are a many-tile problem like prefill: one 256-row activation
  // tile per CTA in the persistent kernel (batch 128: 145 us against 176 us; batch 256: 165 us against 330 us with
  // two 128-row tiles per weight tile).
  if (bn == 128 && ceil_div(N, kBlockW) >= 4 * mtts_num_sms()) bn = 256;
  // The gate/up projection of a decode step at batch 129..256: 96 weight tiles x 2 activation tiles = 192 CTAs = 1.3 waves of
  // the cluster kernel (the SMs that get two CTAs pull 2 MB over the crossbar: 25.4 us per launch inside a 28-layer graph)
  // against 48 CTA-pair tiles of 256 x 256 (every busy SM pulls 1 MB: its 128 weight rows + its 128 activation rows): 18.6 us,
  // decode step at batch 256 4.46 -> 4.33 ms. (Before the SwiGLU epilogue was restructured — swiglu_chunk — the pair route
  // LOST, 32.5 us: its MMAs took 14.7 us and the drain of the 128 x 256 tile 23 us. A stream-K split of the 48 tiles over
  // all 74 pairs with fp32 contributions through an L2 workspace balanced the MMA phase to 10 us but its owner epilogue
  // — contribution loads in front of every chunk — ended at 25.5 us; dropped.) MTTS_GEMM_WIDE_PAIR=0 restores the cluster route.
  static int wide_pair = -1;
  if (wide_pair < 0) {
    const char* e = getenv("MTTS_GEMM_WIDE_PAIR");
    wide_pair = e ? atoi(e) : 1;
  }
  if (wide_pair && bn == 128 && M > 128 && ceil_div(N, 2 * kBlockW) >= mtts_num_sms() / 4 &&
      ceil_div(N, 2 * kBlockW) <= mtts_num_sms() / 2)
    bn = 256;
  const int bk = kSwizzleBytes / eb;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = ceil_div(K, bk);
  const int tiles_n = ceil_div(N, kBlockW), tiles_m = ceil_div(M, bn);
  const int tiles = tiles_n * tiles_m;
  int splits = pick_splits(tiles, p.kb_total, bn, tiles_m);
  p.kb_per_split = ceil_div(p.kb_total, splits);
  while (splits > 1 && (splits - 1) * p.kb_per_split >= p.kb_total) {  // every slice must get >= 1 k-block
    splits /= 2;
    p.kb_per_split = ceil_div(p.kb_total, splits);
  }
  p.splits = splits;
  p.out = out; p.ldo = ldo; p.out_bf16 = out_dtype == MTTS_DTYPE_BF16; p.out_f16 = out_dtype == MTTS_DTYPE_F16;
  p.bias = bias; p.gamma = gamma; p.residual = residual; p.ldr = ldr; p.flags = flags;
  const int oeb = (p.out_bf16 || p.out_f16) ? 2 : 4;
  p.vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) % (4 * oeb)) == 0) &&
             (!(flags & MTTS_EPI_RESIDUAL) || ((ldr % 4 == 0) && (reinterpret_cast<uintptr_t>(residual) % (4 * oeb)) == 0));
  (void)workspace; (void)workspace_bytes;
  CUtensorMap tw, tx;
  int rc = get_tmap(w, N, K, ldw, kBlockW, ecode, &tw);
  if (rc) return rc;
  // large problems on the persistent route: CTA pairs (cta_group::2) on 256 x 256 tiles
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("MTTS_GEMM_2CTA");
    pair_mode = e ? atoi(e) : 1;  // MTTS_GEMM_2CTA=0: the single-CTA persistent kernel
  }
  const bool pair = pair_mode > 0 && bn == 256 && !persist_disabled();
  if (pair) p.pair_tiles_n = ceil_div(N, 2 * kBlockW);
  rc = get_tmap(x, M, K, ldx, pair ? kPBN / 2 : bn, ecode, &tx);
  if (rc) return rc;
  dim3 grid(tiles_n, tiles_m, splits);
  if (in_dtype == MTTS_DTYPE_BF16) return dispatch<bf16>(bn, tw, tx, p, grid, stream);
  if (in_dtype == MTTS_DTYPE_F16) return dispatch<__half>(bn, tw, tx, p, grid, stream);
  return dispatch<float>(bn, tw, tx, p, grid, stream);
}

// ---- fused LM heads: per-quarter best / second-best logit keys instead of logits (called by mtts_heads8_sample) ------------
int mtts_gemm_heads_argmax(const void* x, long long ldx, const void* w, long long ldw, int M, int N, int K, int n_chan,
                           const int* chan_lo, const int* chan_hi, unsigned int* keys, cudaStream_t stream) {
  MTTS_REQUIRE(M > 64 && N % 32 == 0 && n_chan >= 1 && n_chan <= 8, "heads argmax: needs M > 64, N %% 32 == 0");
  MTTS_REQUIRE(!persist_disabled(), "heads argmax: needs the CTA-pair kernel (MTTS_GEMM_NO_PERSIST is set)");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = ceil_div(K, Traits<bf16>::kBlockK);
  p.kb_per_split = p.kb_total;
  p.splits = 1;
  p.out_bf16 = 1;
  p.argmax_keys = keys;
  p.n_quarters = N / 32;
  p.n_chan = n_chan;
  for (int c = 0; c < n_chan; ++c) { p.chan_lo[c] = chan_lo[c]; p.chan_hi[c] = chan_hi[c]; }
  p.pair_tiles_n = ceil_div(N, 2 * kBlockW);
  CUtensorMap tw, tx;
  int rc = get_tmap(w, N, K, ldw, kBlockW, 2, &tw);
  if (rc) return rc;
  rc = get_tmap(x, M, K, ldx, kPBN / 2, 2, &tx);
  if (rc) return rc;
  dim3 grid(ceil_div(N, kBlockW), ceil_div(M, kPBN), 1);
  return dispatch<bf16>(256, tw, tx, p, grid, stream);
}

// ---- split-K with the reduction in the consumer -------------------------------------------------------------------
static int splitk_bn(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

// k-blocks per slice / number of slices: aim at one CTA per SM for the 256-row tile (it owns the SM's shared memory)
// and two for the smaller tiles; every slice keeps >= 2 k-blocks (128 of K).
static void splitk_plan(int M, int N, int K, int* kb_per_split, int* splits) {
  const int bn = splitk_bn(M);
  const int kb_total = ceil_div(K, Traits<bf16>::kBlockK);
  const int tiles = ceil_div(N, kBlockW) * ceil_div(M, bn);
  static int target_env = -1;
  if (target_env < 0) {
    const char* e = getenv("MTTS_SPLITK_TARGET");
    target_env = e ? atoi(e) : 0;
  }
  const int target = target_env > 0 ? target_env : (bn == 256 ? 1 : 2) * mtts_num_sms();
  int s = target / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  int per = ceil_div(kb_total, s);
  if (per < 2) per = kb_total < 2 ? kb_total : 2;
  *kb_per_split = per;
  *splits = ceil_div(kb_total, per);
}

extern "C" int mtts_gemm_splitk_splits(int M, int N, int K) {
  int per, s;
  splitk_plan(M, N, K, &per, &s);
  return s;
}

extern "C" size_t mtts_gemm_splitk_workspace_bytes(int M, int N, int K) {
  return (size_t)mtts_gemm_splitk_splits(M, N, K) * (size_t)M * (size_t)N * sizeof(float);
}

extern "C" int mtts_gemm_splitk(const void* x, long long ldx, const void* w, long long ldw, float* partials,
                                size_t partial_bytes, int M, int N, int K, int* splits_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(M > 0 && M <= 256 && N > 0 && K > 0, "mtts_gemm_splitk: needs 1 <= M <= 256 (got M=%d N=%d K=%d)", M, N, K);
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
               "mtts_gemm_splitk: x and w must be 16-byte aligned");
  MTTS_REQUIRE((ldx * 2) % 16 == 0 && (ldw * 2) % 16 == 0 && ldx >= K && ldw >= K, "mtts_gemm_splitk: bad leading dimensions");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = ceil_div(K, Traits<bf16>::kBlockK);
  splitk_plan(M, N, K, &p.kb_per_split, &p.splits);
  MTTS_REQUIRE(partials != nullptr && partial_bytes >= (size_t)p.splits * M * N * sizeof(float),
               "mtts_gemm_splitk: workspace too small (%zu bytes, need %zu)", partial_bytes,
               (size_t)p.splits * M * N * sizeof(float));
  p.partials = partials;
  if (splits_out) *splits_out = p.splits;
  const int bn = splitk_bn(M);
  CUtensorMap tw, tx;
  int rc = get_tmap(w, N, K, ldw, kBlockW, 2, &tw);
  if (rc) return rc;
  rc = get_tmap(x, M, K, ldx, bn, 2, &tx);
  if (rc) return rc;
  dim3 grid(ceil_div(N, kBlockW), ceil_div(M, bn), p.splits);
  switch (bn) {
    case 16: return launch_partial<16>(tw, tx, p, grid, stream);
    case 32: return launch_partial<32>(tw, tx, p, grid, stream);
    case 64: return launch_partial<64>(tw, tx, p, grid, stream);
    case 128: return launch_partial<128>(tw, tx, p, grid, stream);
    default: return launch_partial<256>(tw, tx, p, grid, stream);
  }
}

#ifdef MTTS_GEMM_TRACE
extern "C" int mtts_debug_gemm_trace(unsigned long long* dst_host, int n) {
  return (int)cudaMemcpyFromSymbol(dst_host, g_gemm_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif
