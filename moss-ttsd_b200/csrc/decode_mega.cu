// Small-batch decode step as ONE persistent kernel (batch <= 4 rows, MOSS-TTSD-v0.5 / Qwen3-1.7B widths).
//
// At batch 1..4 a decode step is a pure weight stream: 3.47 GB of bf16 matrices are read once and every other
// tensor is a few KB. A chain of ~230 kernels (even PDL-chained inside a CUDA graph) leaves the HBM idle at
// every kernel boundary, so here the whole layer stack + LM heads run in one cooperative grid of one CTA per SM:
//
//   * every weight matrix [N, K] is split by OUTPUT ROWS over the CTAs (full K per CTA -> no cross-CTA reduction);
//     a CTA's share of a matrix is one contiguous byte range of the row-major matrix;
//   * one producer warp per CTA walks the CTA's shares of ALL matrices of the step, in order, and streams them
//     with 1-D bulk async copies (TMA engine, mbarrier complete_tx) into a 4-stage shared-memory ring of
//     [16 rows x 1024 k] chunks; it depends on nothing but ring space, so HBM keeps streaming while the consumer
//     warps wait for activations, normalise, or run attention;
//   * 16 consumer warps split a chunk along K (64 k each); weights are the A operand of bf16 mma.sync m16n8k16
//     (ldmatrix from the padded ring rows, conflict-free), the activation rows (n = batch row) are the B operand,
//     read from a shared-memory copy of the phase input; a 16-warp shared-memory reduction finishes a 16-row tile
//     and the epilogue (bf16 rounding / residual / SwiGLU) writes it;
//   * the code is kept SMALL on purpose (one generic projection routine, one staging routine, no per-phase
//     template copies): the consumer path runs once per layer, so a body that overflows the 32 KB instruction cache
//     pays an L2 instruction fetch for every line of every layer (measured: 181 KB of SASS -> 2.5x slower);
//   * there is NO grid barrier: activations travel between phases (q/k/v projection, attention, o_proj, gate/up,
//     down) as 8-byte {payload, tag} words written with one store and polled by their readers (the "LL" protocol
//     of collective libraries): one L2 round trip from producer to consumer, no fences, and a reader only ever
//     waits for the words it needs. The tag encodes (launch, layer, phase), so stale words never match. Every phase
//     consumes the outputs of ALL CTAs of the previous phase, which makes single buffers safe (a writer two phases
//     ahead implies every reader has moved on); the cooperative launch guarantees co-residency of the pollers;
//   * attention at these sizes is latency-, not bandwidth-bound: the first B x Hkv x nsplit CTAs each also own a key
//     range of one (row, kv head) and issue their K/V loads before they start waiting for q.
//
// Rounding points follow the same bf16 eager reference as the kernel-chain path (lm_ops.cu, gemm_tc.cu epilogue,
// attention.cu); see SURVEY.md Appendix B. Replaces, for one decode row per sequence, the Qwen3 layer stack that
// AsteroidTTSInstruct.forward drives (modeling_asteroid.py:252-285) and the 8 lm_heads (:287-288).
#include "common.cuh"
#include "mtts_internal.h"
#include "sm100.cuh"
#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kH = 2048, kI = 6144, kD = 128, kHq = 16, kHkv = 8, kG = 2;
constexpr int kNQKV = (kHq + 2 * kHkv) * kD;  // 4096
constexpr int kCW = 16;                       // consumer warps
constexpr int kConsumers = kCW * 32;          // 512
constexpr int kThreads = kConsumers + 32;     // + 1 producer warp
constexpr int kChunkK = 1024;                 // k elements of one ring chunk
constexpr int kRowBytes = kChunkK * 2;
constexpr int kRowPitch = kRowBytes + 16;     // 16-byte skew per row: ldmatrix phases hit 32 distinct banks
constexpr int kTileRows = 16;
constexpr int kStageBytes = kTileRows * kRowPitch;
// Ring depth. Four stages at batch 2-4: with five the CTA would need the 228 KB shared-memory carve-out, which leaves 28 KB
// of L1 for the stack frames of 544 threads and the read-only loads; four stages fit the 196 KB carve-out (60 KB of L1) and
// the step is 5 % (batch 1) to 9 % (batch 2) faster. Three stages (164 KB) lose again: the ring gets too shallow. At batch 1
// the activation buffer is a quarter of the size, so a fifth stage fits the same 196 KB carve-out.
#ifndef MTTS_MEGA_STAGES
#define MTTS_MEGA_STAGES 4
#endif
#ifndef MTTS_MEGA_STAGES_B1
#define MTTS_MEGA_STAGES_B1 5
#endif
constexpr int kMaxStages = 6;
__host__ __device__ constexpr int ring_stages(int B) { return B == 1 ? MTTS_MEGA_STAGES_B1 : MTTS_MEGA_STAGES; }
constexpr int kMaxB = 4;
constexpr int kActPitch = kI + 8;             // bf16 elements; rows 12304 B apart -> conflict-free B-fragment loads
constexpr int kWsStride = kD + 4;             // attention partial: [M, L, -, -, O[128]] (one LL word per float)
constexpr int kTagsPerLayer = 8;
#ifndef MTTS_MEGA_KV_DEPTH
#define MTTS_MEGA_KV_DEPTH 1
#endif
constexpr int kKvDepth = MTTS_MEGA_KV_DEPTH;  // attention passes (32 keys) in flight per unit

constexpr int kRedTile = kTileRows * kMaxB;   // floats one warp contributes to the tile reduction
constexpr int kMaxCtas = 160;

// Shared-memory layout: [misc | red | act(B) | ring(B)] -- everything the batch-size-independent routines touch sits at a
// fixed offset; only the ring (consume_matrix<B>, the producer) moves with the batch size.
constexpr int kSmemMisc = 4096;                      // barriers, block-reduce scratch, residual rows, RoPE + sentinel tables
constexpr int kSmemRed = 2 * kCW * kRedTile * 4;     // double-buffered 16-warp tile reduction
constexpr int kOffRed = kSmemMisc;
constexpr int kOffAct = kSmemMisc + kSmemRed;        // B activation rows + 256 zero bytes
__host__ __device__ constexpr int act_bytes(int B) { return (B * kActPitch * 2 + 256 + 127) / 128 * 128; }
__host__ __device__ constexpr int ring_offset(int B) { return kOffAct + act_bytes(B); }
__host__ __device__ constexpr int smem_total(int B) { return ring_offset(B) + ring_stages(B) * kStageBytes; }
// the attention scratch (q, k, v of the new token + 16 warps x 2 heads of partials: 18.9 KB) lies over red + act; it ends
// below the zero bytes that follow act row B-1 even at batch 1
constexpr int kAttnScratch = (4 * kD + kCW * kG * kWsStride) * 4;
static_assert(kAttnScratch <= kSmemRed + kActPitch * 2, "attention scratch must stay below the zero pad at batch 1");
static_assert(smem_total(1) <= 196 * 1024 - 2048 && smem_total(kMaxB) <= 196 * 1024 - 2048, "196 KB carve-out");
static_assert(ring_stages(1) <= kMaxStages && ring_stages(kMaxB) <= kMaxStages, "barrier arrays");

struct MegaParams {
  const mtts_lm_layer* layers;
  int num_layers;
  const bf16* heads;
  int vpad;
  const bf16* final_norm;
  const float* inv_freq;
  const int* positions;
  const int* block_table;
  int max_pages, page_shift, num_pages;
  const bf16* x_in;   // [B, H] embedding sum (plain bf16, written by the previous kernel)
  uint2* x_ll[2];     // [B][H/2] words: [0] = layer input / output of down, [1] = output of o_proj
  uint2* qkv_ll;      // [B][NQKV/2]
  uint2* h_ll;        // [B][I/2]
  uint2* ws_ll;       // [B*Hkv*nsplit*G][kWsStride] attention partials, one fp32 per word
  bf16* logits;
  long long ld_logits;
  int nsplit;
  int sentinel;  // 1: one warp watches sentinel words before the CTA-wide verified load
  float eps, scale_log2;
  unsigned int* tag_base;  // [1] device word: tags used by earlier launches
  int* err_flag;
  long long* prof;  // optional: cycles CTA 0 spent per phase, summed over layers (+ per-CTA stamps in trace builds)
};

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// K/V rows are read with an L2 evict_last hint: the weight stream (evict_first) pushes 3.5 GB through L2 every step and
// keeps the DRAM queues full, so a K/V load that misses L2 waits behind ~19 MB of outstanding bulk copies.
#ifndef MTTS_MEGA_KV_EVICT_LAST
#define MTTS_MEGA_KV_EVICT_LAST 1
#endif
__device__ __forceinline__ uint4 ld_kv(const void* p) {
#if MTTS_MEGA_KV_EVICT_LAST
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(kEvictLast));
  return r;
#else
  return ld_nc_v4(p);
#endif
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

// ---- LL words: {payload, tag} in one aligned 8-byte store; a reader polls until the tag is the one it expects.
// (64-bit SCALAR accesses: single-copy atomic in the PTX memory model, unlike the elements of a vector access.)
__device__ __forceinline__ void ll_store(uint2* p, uint32_t data, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)tag << 32) | data) : "memory");
}
// Bounded waits WITHOUT a trap: a wait that exceeds 2^25 polls (>= 20 s at one L2 round trip per poll: a protocol bug,
// or the GPU taken away from this cooperative grid for that long) raises the abort word; every other wait of the launch
// notices it within 1024 polls and gives up too, so the kernel runs to its end (with garbage results), device error
// flag 5 is set and the host raises — the context survives and other streams' work (e.g. an overlapped codec decode)
// is unaffected.
__device__ unsigned int g_mega_abort;   // tag0 + 1 of the launch that gave up (0: none)
__device__ int* g_mega_err;             // the err_flag of the current launch (every CTA writes the same pointer)
__device__ __noinline__ bool ll_slow_check(uint32_t tag, uint32_t spins, uint32_t launch_key) {
  if (*reinterpret_cast<volatile unsigned int*>(&g_mega_abort) == launch_key) return true;
  if (spins < (1u << 25)) return false;
  printf("mtts: decode_mega wait for tag %u timed out (block %d thread %d): aborting the launch\n", tag, blockIdx.x,
         threadIdx.x);
  atomicExch(&g_mega_abort, launch_key);
  if (g_mega_err) *g_mega_err = 5;
  __threadfence();
  return true;
}
// Pollers issue ALL their loads first and compare the tags afterwards (one L2 round trip per attempt, however many
// words a thread needs); on a mismatch the whole batch is re-issued.
__device__ __forceinline__ void ll_issue(const uint2* p, uint32_t& d, uint32_t& t) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  d = (uint32_t)v;
  t = (uint32_t)(v >> 32);
}
__device__ __forceinline__ void ll_issue2(const uint2* p, uint32_t& d0, uint32_t& t0, uint32_t& d1, uint32_t& t1) {
  unsigned long long v0, v1;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(p) : "memory");
  d0 = (uint32_t)v0; t0 = (uint32_t)(v0 >> 32);
  d1 = (uint32_t)v1; t1 = (uint32_t)(v1 >> 32);
}
__shared__ uint32_t s_launch_key;  // tag0 + 1 of this launch
// tags of layer l's words: +1 qkv, +2 attention partials, +3 x after o_proj, +4 h, +5 x after down
__device__ __forceinline__ uint32_t layer_tag(int l) { return s_launch_key - 1u + (uint32_t)l * kTagsPerLayer; }
struct LLSpin {
  uint32_t spins = 0;
  // true: give up (this launch has been aborted)
  __device__ __forceinline__ bool miss(uint32_t tag) {
    if ((++spins & 1023u) != 0) return false;
    return ll_slow_check(tag, spins, s_launch_key);
  }
};

// Rows [r0, r1) of an N-row matrix owned by CTA `cta`: units of `unit` rows dealt out evenly, the remainder rotated
// by `rot` so that the longer shares of the four per-layer matrices land on different CTAs.
struct Slice { int r0, r1; };
__device__ __forceinline__ Slice slice_rows(int n_rows, int unit, int rot, int cta, int G) {
  const int units = n_rows / unit;
  const int c = (int)(((unsigned)cta + (unsigned)rot) % (unsigned)G);
  const int base = units / G, rem = units % G;
  const int u0 = c * base + min(c, rem);
  const int nu = base + (c < rem ? 1 : 0);
  return Slice{u0 * unit, (u0 + nu) * unit};
}
// first row of the last 16-row tile of a share (its words are the last a producer publishes in a phase)
__device__ __forceinline__ int last_tile_row(Slice s) { return s.r0 + ((s.r1 - s.r0 - 1) / kTileRows) * kTileRows; }

// The matrices of the step in consumption order: m = 0 q/k/v, 1 o_proj, 2 gate/up, 3 down (per layer), 4 LM heads.
// Shares are the same in every layer (rotation depends on m only), so they and the sentinel addresses derived from
// them are computed once per launch.
__host__ __device__ constexpr int mat_rows(int m) { return m == 0 ? kNQKV : (m == 2 ? 2 * kI : kH); }
__host__ __device__ constexpr int mat_unit(int m) { return m == 2 ? 4 : 2; }
#ifndef MTTS_MEGA_ROT
#define MTTS_MEGA_ROT 0
#endif
__host__ __device__ constexpr int mat_rot(int m) { return m * MTTS_MEGA_ROT; }

// Polling by all 512 consumer threads of all CTAs would swamp the L2 request queues that the weight stream also
// needs, so before a CTA-wide verified load ONE warp watches a sentinel word per producer CTA (the first word of the
// producer's last tile, batch row B-1); the other warps sleep in the hardware barrier that follows. The sentinel is
// only a hint: the load after it still checks every tag and retries.
// `offs`: shared-memory table of the sentinel word offsets (one per producer CTA), built once per launch.
__device__ __noinline__ void sentinel_wait(const uint2* base, const int* offs, int n, uint32_t tag) {
  const int lane = threadIdx.x & 31;
  LLSpin sp;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int c = lane + 32 * i;
      uint32_t d, t = tag;
      if (c < n) ll_issue(base + offs[c], d, t);
      ok = ok && (t == tag);
    }
    if (__all_sync(0xffffffffu, ok)) break;
    if (sp.miss(tag)) break;
  }
}

// ---- shared-memory layout. The phase routines are not inlined (the step's code must stay within the instruction
// cache), and anything handed to them by reference or beyond the register-passed arguments travels through the LOCAL
// stack frame: 544 threads' frames do not fit the 60 KB of L1 the ring leaves over, and an LDL that misses sits at the
// head of every phase. So the routines take two or three scalar arguments and find everything else here.
extern __shared__ __align__(128) uint8_t mega_smem[];
__shared__ MegaParams s_params;
__shared__ Slice s_sl[5];  // this CTA's shares of the four per-layer matrices and of the LM heads
// Ring chunks / 16-row tiles this CTA consumes before matrix m of a layer ([4] = per layer): the consumer's ring
// position and reduction buffer follow from (layer, matrix), so no loop-carried state has to survive the phase calls
// (values that live across a call are spilled to the stack frame: see above).
__shared__ int s_cpre[5], s_tpre[5];
__shared__ int s_unit;  // attention unit of this CTA, or -1
__device__ __forceinline__ bf16* sm_act() { return reinterpret_cast<bf16*>(mega_smem + kOffAct); }
__device__ __forceinline__ float* sm_red() { return reinterpret_cast<float*>(mega_smem + kOffRed); }
__device__ __forceinline__ uint64_t* sm_full() {
  return reinterpret_cast<uint64_t*>(mega_smem);
}
__device__ __forceinline__ uint64_t* sm_empty() { return sm_full() + kMaxStages; }
__device__ __forceinline__ float* sm_scratch() { return reinterpret_cast<float*>(sm_empty() + kMaxStages + 2); }  // [16][kMaxB]
__device__ __forceinline__ float* sm_res(int which) {  // [kMaxB][16] raw residual values of this CTA's o_proj (0) / down (1) rows
  return sm_scratch() + kCW * kMaxB + which * (kMaxB * 16);
}
__device__ __forceinline__ float* sm_rope() { return sm_scratch() + kCW * kMaxB + 2 * (kMaxB * 16); }  // [128]
// sentinel word offsets per producer CTA: 0 = o_proj output, 1 = down output, 2 = SwiGLU output
__device__ __forceinline__ int* sm_sent(int which) { return reinterpret_cast<int*>(sm_rope() + 128) + which * kMaxCtas; }

// opaque identity: keeps the compiler from proving that a phase routine returns its argument (it would then re-read the
// caller's stack copy instead of using the returned register)
__device__ __forceinline__ int launder(int v) {
  asm volatile("" : "+r"(v));
  return v;
}

struct Ring {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
  uint32_t it;
};

// ------------------------------------------------------------------ consumer: one matrix share
enum { EPI_QKV = 0, EPI_WO = 1, EPI_GU = 2, EPI_WD = 3, EPI_HEADS = 4 };

// Profiling counters live in shared memory while the kernel runs (a global read-modify-write per tick would itself
// cost ~0.3 us on the critical path of CTA 0) and are flushed once at the end.
__shared__ long long s_prof[32];

// act: [kMaxB][kActPitch] bf16 phase input; `out`: LL destination (qkv / x / h words, `out_wpr` words per batch row) or
// unused for the heads; `res`: [kMaxB][16] raw residual values of the rows of this share (o_proj / down only);
// `tag`: tag of the words this phase publishes.
// Bounded wait with the bookkeeping out of line: the fast path is one try_wait and one branch.
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// act: [kMaxB][kActPitch] bf16 phase input, followed by 256 zero bytes (the B operand of the unused batch columns);
// `out`: LL destination (qkv / x / h words, `out_wpr` words per batch row) or unused for the heads; `res`: [kMaxB][16]
// raw residual values of the rows of this share (o_proj / down only); `tag`: tag of the words this phase publishes.
// The loop is issue-bound if it is not kept lean (4 warps per scheduler, 64 KB of weights per trip): no modulo, no
// select, no bookkeeping inside.
template <int kB>
__device__ __noinline__ int consume_matrix(int l, int epi) {
  const MegaParams& p = s_params;
  const Slice s = s_sl[epi];
  const int kch = epi == EPI_WD ? kI / kChunkK : kH / kChunkK;
  const bf16* act = sm_act();
  float* red = sm_red();
  const int m4 = epi & 3;  // the LM heads (epi 4) come first in "layer" num_layers
  int red_buf = (l * s_tpre[4] + s_tpre[m4]) & 1;
  constexpr int kStages = ring_stages(kB);
  Ring rg{mega_smem + ring_offset(kB), sm_full(), sm_empty(), (uint32_t)(l * s_cpre[4] + s_cpre[m4])};
  const uint32_t tag = layer_tag(l) + (epi == EPI_QKV ? 1u : (epi == EPI_WO ? 3u : (epi == EPI_GU ? 4u : 5u)));
  uint2* out = epi == EPI_QKV ? p.qkv_ll : (epi == EPI_WO ? p.x_ll[1] : (epi == EPI_GU ? p.h_ll : p.x_ll[0]));
  const int out_wpr = epi == EPI_QKV ? kNQKV / 2 : (epi == EPI_GU ? kI / 2 : kH / 2);
  const float* res = sm_res(epi == EPI_WD ? 1 : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t a_off = (uint32_t)((lane & 15) * kRowPitch + (warp * 64 + (lane >> 4) * 8) * 2);
  const bf16* b_src = g < kB ? act + g * kActPitch + warp * 64 + 2 * t : act + kB * kActPitch + 2 * t;
  const int kmul = g < kB ? kChunkK : 0;
  uint32_t st = rg.it % kStages, par = (rg.it / kStages) & 1;
  // B fragments (k16 x n8, "col") of this warp's 64-wide k slice of two chunks: batch row g, k = 2t,2t+1 and +8
  uint32_t bq[8][2];
  auto load_bq = [&](int kg) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bq[i][0] = *reinterpret_cast<const uint32_t*>(b_src + (kg + (i >> 2)) * kmul + (i & 3) * 16);
      bq[i][1] = *reinterpret_cast<const uint32_t*>(b_src + (kg + (i >> 2)) * kmul + (i & 3) * 16 + 8);
    }
  };
  if (kch == 2) load_bq(0);
#pragma unroll 1
  for (int r = s.r0; r < s.r1; r += kTileRows) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int kg = 0; kg < kch; kg += 2) {  // two chunks per trip (kch is 2 or 6)
      if (kch != 2) load_bq(kg);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        mbar_wait_lean(&rg.full[st], par);
        const uint32_t a_base = smem_u32(rg.stages + st * kStageBytes) + a_off;
        uint32_t a[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(a[ks], a_base + ks * 32);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_16x8x16(c, a[ks], bq[kc * 4 + ks][0], bq[kc * 4 + ks][1]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty[st]);
        if (++st == kStages) { st = 0; par ^= 1; }
      }
    }
    // 16-warp reduction of the [16 rows x 4 batch] tile (accumulator columns 2t, 2t+1; only t < 2 are real rows)
    float* rb = red + red_buf * (kCW * kRedTile) + warp * kRedTile;
    if (t < 2) {
      *reinterpret_cast<float2*>(rb + g * kMaxB + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(rb + (g + 8) * kMaxB + 2 * t) = make_float2(c[2], c[3]);
    }
    consumer_sync();
    if (tid < kRedTile) {  // two warps: lane + 4 holds the next row of the same batch column
      const float* rr = red + red_buf * (kCW * kRedTile) + tid;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kCW; ++w) v += rr[w * kRedTile];
      const int row = r + (tid >> 2), n = tid & 3;
      const bool ok = row < s.r1 && n < kB;
      if (epi == EPI_HEADS) {
        if (ok) p.logits[(size_t)n * p.ld_logits + row] = __float2bfloat16_rn(v);
      } else if (epi == EPI_GU) {
        // rows 2j = gate_j, 2j+1 = up_j: bf16(gate), bf16(silu), bf16(up), bf16(product)
        const float up = __shfl_down_sync(0xffffffffu, v, 4);
        const float hv = bf16_round(silu_f(bf16_round(v))) * bf16_round(up);
        const float hn = __shfl_down_sync(0xffffffffu, hv, 8);
        if (ok && !(row & 3)) ll_store(out + (size_t)n * out_wpr + (row >> 2), pack_bf16(hv, hn), tag);
      } else {
        float xv = v;  // q/k/v: bf16(linear);   o_proj / down: hidden = residual + bf16(linear), both bf16 tensors
        if (epi != EPI_QKV && ok) xv = bf16_round(res[n * 16 + (row - s.r0)] + bf16_round(v));
        const float nx = __shfl_down_sync(0xffffffffu, xv, 4);
        if (ok && !(row & 1)) ll_store(out + (size_t)n * out_wpr + (row >> 1), pack_bf16(xv, nx), tag);
      }
    }
    red_buf ^= 1;
  }
  return launder(l);  // handed back so that the caller need not keep it across the call (it would live in the stack frame)
}

// Stage B rows of 2048 elements (LL words, or the plain embedding sum) into act with RMSNorm on the way (every CTA
// does it redundantly: 8 KB per row from L2):
//   v = mean(x^2); y = bf16(x * rsqrt(v + eps)); out = bf16(w * y)        (lm_ops.cu rmsnorm_kernel)
// The raw values of rows [keep.r0, keep.r1) are kept in `res` for the residual add of the following projection.
// `which`: 0 = layer input (x after down / the embedding sum) -> ln1, residual rows of the o_proj share kept;
//          1 = x after o_proj -> ln2, residual rows of the down share kept;   2 = final norm (nothing kept)
template <int kB>
__device__ __noinline__ int stage_norm(int which, int l) {
  const MegaParams& p = s_params;
  const uint32_t tag = which == 1 ? layer_tag(l) + 3u : layer_tag(l) - kTagsPerLayer + 5u;
  const bf16* __restrict__ w = which == 2 ? p.final_norm
                                          : reinterpret_cast<const bf16*>(which == 0 ? p.layers[l].ln1 : p.layers[l].ln2);
  const bf16* x_plain = (which == 0 && l == 0) ? p.x_in : nullptr;
  const uint2* src_ll = p.x_ll[which == 1 ? 1 : 0];
  const int* sent = sm_sent(which == 1 ? 0 : 1);
  const int n_gemv = (int)gridDim.x;
  bf16* act = sm_act();
  float* scratch = sm_scratch();
  const Slice keep = which == 2 ? Slice{0, 0} : s_sl[which == 0 ? 1 : 3];
  float* res = sm_res(which == 1 ? 1 : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint2 xv[kB];
  const uint2 wv = *reinterpret_cast<const uint2*>(w + tid * 4);
  if (x_plain) {
#pragma unroll
    for (int b = 0; b < kB; ++b) xv[b] = *reinterpret_cast<const uint2*>(x_plain + (size_t)b * kH + tid * 4);
  } else {
    if (p.sentinel) {
      if (warp == 0) sentinel_wait(src_ll + (size_t)(kB - 1) * (kH / 2), sent, n_gemv, tag);
      consumer_sync();
    }
    LLSpin sp;
    while (true) {
      bool ok = true;
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        uint32_t t0, t1;
        ll_issue2(src_ll + (size_t)b * (kH / 2) + tid * 2, xv[b].x, t0, xv[b].y, t1);
        ok = ok && t0 == tag && t1 == tag;
      }
      if (ok) break;
      if (sp.miss(tag)) break;
    }
  }
  const int e = tid * 4;
  const bool keeps = e + 3 >= keep.r0 && e < keep.r1;
#pragma unroll
  for (int b = 0; b < kB; ++b) {
    const float a[4] = {bf16lo(xv[b].x), bf16hi(xv[b].x), bf16lo(xv[b].y), bf16hi(xv[b].y)};
    const float ss = warp_sum(fmaf(a[0], a[0], fmaf(a[1], a[1], fmaf(a[2], a[2], a[3] * a[3]))));
    if (lane == 0) scratch[warp * kMaxB + b] = ss;
    if (keeps) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e + j >= keep.r0 && e + j < keep.r1) res[b * 16 + (e + j - keep.r0)] = a[j];
    }
  }
  consumer_sync();
#pragma unroll
  for (int b = 0; b < kB; ++b) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kCW; ++i) tot += scratch[i * kMaxB + b];
    const float inv = rsqrtf(tot * (1.0f / kH) + p.eps);
    const float o0 = bf16_round(bf16lo(xv[b].x) * inv) * bf16lo(wv.x);
    const float o1 = bf16_round(bf16hi(xv[b].x) * inv) * bf16hi(wv.x);
    const float o2 = bf16_round(bf16lo(xv[b].y) * inv) * bf16lo(wv.y);
    const float o3 = bf16_round(bf16hi(xv[b].y) * inv) * bf16hi(wv.y);
    *reinterpret_cast<uint2*>(act + b * kActPitch + tid * 4) = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
  }
  consumer_sync();
  return launder(l);  // handed back so that the caller need not keep it across the call (it would live in the stack frame)
}

// Stage the B rows of the SwiGLU output (6144 elements each) into act, unchanged.
template <int kB>
__device__ __noinline__ int stage_h(int l) {
  const MegaParams& p = s_params;
  const uint32_t tag = layer_tag(l) + 4u;
  const int* sent = sm_sent(2);
  const int n_gemv = (int)gridDim.x;
  bf16* act = sm_act();
  const int tid = threadIdx.x, warp = tid >> 5;
  if (p.sentinel) {
    if (warp == 0) sentinel_wait(p.h_ll + (size_t)(kB - 1) * (kI / 2), sent, n_gemv, tag);
    consumer_sync();
  }
  uint2 xv[3][kB];
  LLSpin sp;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int pc = 0; pc < 3; ++pc)
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        uint32_t t0, t1;
        ll_issue2(p.h_ll + (size_t)b * (kI / 2) + pc * (kH / 2) + tid * 2, xv[pc][b].x, t0, xv[pc][b].y, t1);
        ok = ok && t0 == tag && t1 == tag;
      }
    if (ok) break;
    if (sp.miss(tag)) break;
  }
#pragma unroll
  for (int pc = 0; pc < 3; ++pc)
#pragma unroll
    for (int b = 0; b < kB; ++b) *reinterpret_cast<uint2*>(act + b * kActPitch + pc * kH + tid * 4) = xv[pc][b];
  consumer_sync();
  return launder(l);  // handed back so that the caller need not keep it across the call (it would live in the stack frame)
}

// One polling batch of the split-KV merge: M, L and four O values of up to four splits of one (row, head); absent splits
// come back as M = -inf. The spin is bounded without a call (see stage_attn_out); true = gave up.
__device__ __forceinline__ bool poll_partials(const uint2* base, int s0, int nsplit, int d4, uint32_t tag, float (&ms)[4],
                                              float (&ls)[4], float4 (&ov)[4]) {
  uint32_t spins = 0;
  while (true) {
    uint32_t bad = 0u;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      uint32_t d0 = 0xff800000u, d1 = 0u, d2 = 0u, d3 = 0u, d4v = 0u, d5 = 0u;
      if (s0 + s < nsplit) {
        const uint2* w0 = base + (size_t)(s0 + s) * kG * kWsStride;
        uint32_t t0, t1, t2, t3, t4, t5;
        ll_issue2(w0, d0, t0, d1, t1);
        ll_issue2(w0 + 4 + d4, d2, t2, d3, t3);
        ll_issue2(w0 + 6 + d4, d4v, t4, d5, t5);
        bad |= (t0 ^ tag) | (t1 ^ tag) | (t2 ^ tag) | (t3 ^ tag) | (t4 ^ tag) | (t5 ^ tag);
      }
      ms[s] = __uint_as_float(d0);
      ls[s] = __uint_as_float(d1);
      ov[s] = make_float4(__uint_as_float(d2), __uint_as_float(d3), __uint_as_float(d4v), __uint_as_float(d5));
    }
    if (bad == 0u) return false;
    if (++spins > (1u << 21)) return true;  // ~2 s of L2 round trips
    if ((spins & 1023u) == 0u && *reinterpret_cast<volatile unsigned int*>(&g_mega_abort) == s_launch_key) return true;
  }
}

// Merge the split-KV partials of the attention phase into act[b][head*128 + d] (bf16, as the o_proj input).
__shared__ int s_layer_keep;
template <int kB>
__device__ __noinline__ int stage_attn_out(int l) {
  const MegaParams& p = s_params;
  const uint32_t tag = layer_tag(l) + 2u;
  // this routine is short of registers at batch 1: the layer index waits in shared memory (not in the stack frame) for
  // the return; written before the first barrier below, read behind the last one
  if (threadIdx.x == 0) s_layer_keep = l;
  bf16* act = sm_act();
  const int units = kB * kHkv * p.nsplit;
  if (p.sentinel && (threadIdx.x >> 5) == 0) {  // sentinel: the last O word of head g = 1 of every unit
    const int lane = threadIdx.x & 31;
    uint32_t spins = 0;
    while (true) {  // (call-free, like the polls below; the verified loads that follow report a real time-out)
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int u = lane + 32 * i;
        uint32_t d, t = tag;
        if (u < units) ll_issue(p.ws_ll + ((size_t)u * kG + 1) * kWsStride + 4 + 127, d, t);
        ok = ok && (t == tag);
      }
      if (__all_sync(0xffffffffu, ok)) break;
      if (++spins > (1u << 21)) break;
      if ((spins & 1023u) == 0u && *reinterpret_cast<volatile unsigned int*>(&g_mega_abort) == s_launch_key) break;
    }
  }
  consumer_sync();
  bool timed_out = false;
#pragma unroll 1
  for (int idx = threadIdx.x; idx < kB * (kH / 4); idx += kConsumers) {
    const int b = idx / (kH / 4), rem = idx % (kH / 4);
    const int head = rem >> 5, d4 = (rem & 31) * 4;
    const int hk = head / kG, g = head % kG;
    const uint2* base = p.ws_ll + ((size_t)((b * kHkv + hk) * p.nsplit) * kG + g) * kWsStride;
    // Splits are merged four at a time (one polling batch each), groups with a running online-softmax state. The first
    // group is peeled and the poll loops contain no call (a call in a poll loop makes the compiler park everything that
    // is live around it in the stack frame -- six accumulators and the layer index, reloaded on the critical path).
    float ms[4], ls[4];
    float4 ov[4];
    timed_out |= poll_partials(base, 0, p.nsplit, d4, tag, ms, ls, ov);
    float M = fmaxf(fmaxf(fmaxf(-1e30f, ms[0]), fmaxf(ms[1], ms[2])), ms[3]);
    float L = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float w = ex2(ms[s] - M);
      L = fmaf(ls[s], w, L);
      o[0] = fmaf(ov[s].x, w, o[0]); o[1] = fmaf(ov[s].y, w, o[1]); o[2] = fmaf(ov[s].z, w, o[2]); o[3] = fmaf(ov[s].w, w, o[3]);
    }
#pragma unroll 1
    for (int s0 = 4; s0 < p.nsplit; s0 += 4) {
      timed_out |= poll_partials(base, s0, p.nsplit, d4, tag, ms, ls, ov);
      float Mn = M;
#pragma unroll
      for (int s = 0; s < 4; ++s) Mn = fmaxf(Mn, ms[s]);
      const float c = ex2(M - Mn);
      M = Mn;
      L *= c;
      o[0] *= c; o[1] *= c; o[2] *= c; o[3] *= c;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float w = ex2(ms[s] - M);
        L = fmaf(ls[s], w, L);
        o[0] = fmaf(ov[s].x, w, o[0]); o[1] = fmaf(ov[s].y, w, o[1]); o[2] = fmaf(ov[s].z, w, o[2]); o[3] = fmaf(ov[s].w, w, o[3]);
      }
    }
    const float inv = 1.0f / L;
    *reinterpret_cast<uint2*>(act + b * kActPitch + head * kD + d4) =
        make_uint2(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv));
  }
  if (timed_out) ll_slow_check(tag, 1u << 25, s_launch_key);  // report + abort flag, with nothing live around the call
  consumer_sync();
  return launder(s_layer_keep);  // handed back so that the caller need not keep it across the call (it would live in the stack frame)
}

// ------------------------------------------------------------------ attention CTAs
// Unit = (batch row, kv head, split), one per attention CTA for the whole step. Per layer: all 32 half-warps issue the
// K/V loads of their first cached key (they do not depend on this step), warps 0..3 wait for q (2 heads), k and v of
// the new token and finish them (per-head RMSNorm + RoPE, lm_ops.cu qknorm_rope_kv_kernel; the owner split appends
// k/v to the cache), the cached keys are walked with an fp32 online softmax (attention.cu; 16 lanes per key row,
// 32 rows per pass, next pass in flight), the owner folds in the new key from shared memory, and the CTA publishes
// one (M, L, O[128]) partial per q head. The loops are deliberately rolled: this code must stay small.
// `rope`: [cos[64] | sin[64]] of this unit's position, bf16-rounded, computed once per step.
struct AttnUnit {
  int b, hk, pos, k_begin, k_end, owner;  // k_end excludes the new key
};
struct AttnLayer { const bf16 *k_pool, *v_pool, *q_norm, *k_norm; };
// Per-CTA constants of the attention unit and the current layer's pointers live in SHARED memory, not in the kernel's
// stack frame: a struct handed to a non-inlined phase routine by reference is read back with LDL, the 544 threads' frames
// do not fit the 60 KB of L1 the ring leaves over, and an LDL that misses L1 sits in front of every K/V address (measured:
// ~1400 cycles per 32-key pass, which deeper K/V prefetching did not shorten).
__shared__ AttnUnit s_un;
__shared__ AttnLayer s_al;

// one key row (this lane's 8 dims of K and V) folded into the online-softmax state of both q heads. q (normed, roped,
// pre-scaled) is re-read from shared memory in every pass: holding it in 16 registers next to the accumulators leaves no
// room under the 96-register cap (17 warps) for a second K/V pass in flight.
__device__ __forceinline__ void attn_fold(const float* s_qs, uint4 kq, uint4 vq, bool valid, unsigned mask,
                                       float (&m)[kG], float (&l)[kG], float (&acc)[kG][8]) {
  float kf[8], vf[8];
  unpack8(kq, kf);
  unpack8(vq, vf);
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const float4 q0 = *reinterpret_cast<const float4*>(s_qs + g * kD);
    const float4 q1 = *reinterpret_cast<const float4*>(s_qs + g * kD + 4);
    float s = q0.x * kf[0];
    s = fmaf(q0.y, kf[1], s); s = fmaf(q0.z, kf[2], s); s = fmaf(q0.w, kf[3], s);
    s = fmaf(q1.x, kf[4], s); s = fmaf(q1.y, kf[5], s); s = fmaf(q1.z, kf[6], s); s = fmaf(q1.w, kf[7], s);
    s += __shfl_xor_sync(mask, s, 8);
    s += __shfl_xor_sync(mask, s, 4);
    s += __shfl_xor_sync(mask, s, 2);
    s += __shfl_xor_sync(mask, s, 1);
    if (!valid) s = -INFINITY;
    const float mx = fmaxf(m[g], s);
    const float corr = ex2(m[g] - mx), pw = ex2(s - mx);
    m[g] = mx;
    l[g] = fmaf(l[g], corr, pw);
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[g][d] = fmaf(pw, vf[d], acc[g][d] * corr);
  }
}

__device__ __noinline__ int attention_layer(int layer) {
  const MegaParams& p = s_params;
  const int unit = s_unit;
  const uint32_t tag_in = layer_tag(layer) + 1u, tag_out = layer_tag(layer) + 2u;
  float* scr = sm_red();  // red + act: both idle between the q/k/v projection and the partial merge
  const float* rope = sm_rope();
  const AttnUnit un = s_un;
  const AttnLayer L = s_al;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // The partials come first (over `red`, whose last readers -- warps 0 and 1 in the epilogue of the q/k/v projection --
  // are long past it when the partials are written behind this routine's first barrier); q, k, v are written at once by
  // warps 0..3 and therefore lie in the act part, which nobody reads after the projection's last tile barrier.
  float* s_part = scr;                            // [16 warps][2][kWsStride]
  float* s_q = scr + kCW * kG * kWsStride;        // [2][128] normed + roped, pre-scaled q (bf16-rounded values)
  float* s_k = s_q + 2 * kD;                      // [128]
  float* s_v = s_q + 3 * kD;                      // [128]
  static_assert(kCW * kG * kWsStride * 4 >= kSmemRed, "q / k / v must lie beyond the reduction buffers");
  const int page_mask = (1 << p.page_shift) - 1;
  const int b = un.b, hk = un.hk, pos = un.pos, k_end = un.k_end;
  const long long head_off = ((long long)hk << p.page_shift) * kD;
  const long long page_stride = ((long long)kHkv << p.page_shift) * kD;
  const int l16 = lane & 15, rgp = warp * 2 + (lane >> 4);
#ifdef MTTS_MEGA_PROFILE
  const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long tprev = prof ? clock64() : 0;
#define ATTN_TICK(slot)                     \
  if (prof) {                               \
    const long long tn = clock64();         \
    s_prof[slot] += tn - tprev;             \
    tprev = tn;                             \
  }
#else
#define ATTN_TICK(slot)
#endif

  auto kv_offset = [&](int key) {
    key = min(key, k_end - 1);  // clamp: the load is always legal, the score is masked
    const int lp = key >> p.page_shift;
    const int page = p.block_table ? __ldg(p.block_table + (long long)b * p.max_pages + lp) : b * p.max_pages + lp;
    return (long long)page * page_stride + head_off + (long long)(key & page_mask) * kD + l16 * 8;
  };
  // The first kKvDepth passes (32 keys each) are requested before the unit starts waiting for q: a pass costs one
  // DRAM / L2 round trip (~1400 cycles under the weight stream), so with one pass in flight the key loop is a chain
  // of round trips (measured: 5850 cycles for 4 passes); with all of them in flight it is one.
  uint4 kb[kKvDepth], vb[kKvDepth];
#pragma unroll
  for (int i = 0; i < kKvDepth; ++i) {
    kb[i] = make_uint4(0, 0, 0, 0);
    vb[i] = kb[i];
    if (un.k_begin + 32 * i < k_end) {
      const long long off = kv_offset(un.k_begin + 32 * i + rgp);
      kb[i] = ld_kv(L.k_pool + off);
      vb[i] = ld_kv(L.v_pool + off);
    }
  }
  ATTN_TICK(16)

  if (warp < 4) {
    const int col = warp < 2 ? (hk * kG + warp) * kD : (warp == 2 ? kHq * kD + hk * kD : (kHq + kHkv) * kD + hk * kD);
    bf16* dst = nullptr;
    if (warp >= 2 && un.owner) {
      const int lp = pos >> p.page_shift;
      int page = -1;
      if (lp < p.max_pages) page = p.block_table ? p.block_table[(long long)b * p.max_pages + lp] : b * p.max_pages + lp;
      if (page < 0 || page >= p.num_pages) {
        if (lane == 0 && p.err_flag) *p.err_flag = 2;
      } else {
        dst = const_cast<bf16*>(warp == 2 ? L.k_pool : L.v_pool) + (long long)page * page_stride + head_off +
              (long long)(pos & page_mask) * kD;
      }
    }
    const bf16* nw = warp < 2 ? L.q_norm : L.k_norm;
    const uint32_t wa = *reinterpret_cast<const uint32_t*>(nw + 2 * lane);
    const uint32_t wb = *reinterpret_cast<const uint32_t*>(nw + 64 + 2 * lane);
    const uint2* src = p.qkv_ll + (size_t)b * (kNQKV / 2) + col / 2;
    uint32_t a, c;  // elements (2l, 2l+1) and (64+2l, 65+2l)
    {
      LLSpin sp;
      while (true) {
        uint32_t ta, tc;
        ll_issue(src + lane, a, ta);
        ll_issue(src + 32 + lane, c, tc);
        if (ta == tag_in && tc == tag_in) break;
        if (sp.miss(tag_in)) break;
      }
    }
    ATTN_TICK(17)
    float x0 = bf16lo(a), x1 = bf16hi(a), x2 = bf16lo(c), x3 = bf16hi(c);
    float* so = warp < 2 ? s_q + warp * kD : (warp == 2 ? s_k : s_v);
    if (warp != 3) {
      float ss = x0 * x0;
      ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
      ss = warp_sum(ss);
      const float inv = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
      x0 = bf16_round(bf16lo(wa) * bf16_round(x0 * inv));
      x1 = bf16_round(bf16hi(wa) * bf16_round(x1 * inv));
      x2 = bf16_round(bf16lo(wb) * bf16_round(x2 * inv));
      x3 = bf16_round(bf16hi(wb) * bf16_round(x3 * inv));
      const float c0 = rope[2 * lane], c1 = rope[2 * lane + 1], s0 = rope[64 + 2 * lane], s1 = rope[65 + 2 * lane];
      const float o0 = bf16_round(bf16_round(x0 * c0) + bf16_round(-x2 * s0));
      const float o1 = bf16_round(bf16_round(x1 * c1) + bf16_round(-x3 * s1));
      const float o2 = bf16_round(bf16_round(x2 * c0) + bf16_round(x0 * s0));
      const float o3 = bf16_round(bf16_round(x3 * c1) + bf16_round(x1 * s1));
      x0 = o0; x1 = o1; x2 = o2; x3 = o3;
    }
    const float qs = warp < 2 ? p.scale_log2 : 1.0f;  // q is stored pre-scaled (softmax scale * log2 e)
    so[2 * lane] = x0 * qs; so[2 * lane + 1] = x1 * qs; so[64 + 2 * lane] = x2 * qs; so[65 + 2 * lane] = x3 * qs;
    if (dst) {
      *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pack_bf16(x0, x1);
      *reinterpret_cast<uint32_t*>(dst + 64 + 2 * lane) = pack_bf16(x2, x3);
    }
  }
  ATTN_TICK(18)
  consumer_sync();
  ATTN_TICK(19)

  const float* s_qs = s_q + l16 * 8;  // this lane's 8 dims of both (pre-scaled) q heads
  float m[kG], l[kG], acc[kG][8];
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    m[g] = -1e30f; l[g] = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[g][d] = 0.f;
  }
#pragma unroll 1
  for (int key = un.k_begin + rgp; key - rgp < k_end; key += 32) {
    const uint4 kq = kb[0], vq = vb[0];
#pragma unroll
    for (int i = 0; i + 1 < kKvDepth; ++i) {
      kb[i] = kb[i + 1];
      vb[i] = vb[i + 1];
    }
    if (key - rgp + 32 * kKvDepth < k_end) {  // keep kKvDepth passes in flight
      const long long off = kv_offset(key + 32 * kKvDepth);
      kb[kKvDepth - 1] = ld_kv(L.k_pool + off);
      vb[kKvDepth - 1] = ld_kv(L.v_pool + off);
    }
    attn_fold(s_qs, kq, vq, key < k_end, 0xffffffffu, m, l, acc);
  }
  if (un.owner && rgp == 0) {  // the new token's key/value (never read back from global memory)
    const uint4 kn = make_uint4(pack_bf16(s_k[l16 * 8], s_k[l16 * 8 + 1]), pack_bf16(s_k[l16 * 8 + 2], s_k[l16 * 8 + 3]),
                                pack_bf16(s_k[l16 * 8 + 4], s_k[l16 * 8 + 5]), pack_bf16(s_k[l16 * 8 + 6], s_k[l16 * 8 + 7]));
    const uint4 vn = make_uint4(pack_bf16(s_v[l16 * 8], s_v[l16 * 8 + 1]), pack_bf16(s_v[l16 * 8 + 2], s_v[l16 * 8 + 3]),
                                pack_bf16(s_v[l16 * 8 + 4], s_v[l16 * 8 + 5]), pack_bf16(s_v[l16 * 8 + 6], s_v[l16 * 8 + 7]));
    attn_fold(s_qs, kn, vn, true, 0x0000ffffu, m, l, acc);  // s_k / s_v hold bf16-rounded values: the packing is exact
  }
  __syncwarp();
  ATTN_TICK(20)
  // merge the two half-warps, then the 16 warps
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const float mo = __shfl_xor_sync(0xffffffffu, m[g], 16);
    const float lo = __shfl_xor_sync(0xffffffffu, l[g], 16);
    const float M = fmaxf(m[g], mo);
    const float wa = ex2(m[g] - M), wb = ex2(mo - M);
    l[g] = l[g] * wa + lo * wb;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[g][d], 16);
      acc[g][d] = acc[g][d] * wa + ao * wb;
    }
    m[g] = M;
    if (lane < 16) {
      float* o = s_part + (warp * kG + g) * kWsStride;
      if (lane == 0) { o[0] = m[g]; o[1] = l[g]; }
#pragma unroll
      for (int d = 0; d < 8; ++d) o[4 + l16 * 8 + d] = acc[g][d];
    }
  }
  consumer_sync();
  if (tid < kG * kD) {
    const int g = tid / kD, d = tid % kD;
    float M = -1e30f;
#pragma unroll 4
    for (int w = 0; w < kCW; ++w) M = fmaxf(M, s_part[(w * kG + g) * kWsStride]);
    float Ls = 0.f, O = 0.f;
#pragma unroll 2
    for (int w = 0; w < kCW; ++w) {
      const float* o = s_part + (w * kG + g) * kWsStride;
      const float wt = ex2(o[0] - M);
      Ls = fmaf(o[1], wt, Ls);
      O = fmaf(o[4 + d], wt, O);
    }
    uint2* w0 = p.ws_ll + ((size_t)(unit * kG + g)) * kWsStride;
    ll_store(w0 + 4 + d, __float_as_uint(O), tag_out);
    if (d == 0) {
      ll_store(w0, __float_as_uint(M), tag_out);
      ll_store(w0 + 1, __float_as_uint(Ls), tag_out);
    }
  }
  consumer_sync();
  ATTN_TICK(21)
#undef ATTN_TICK
  return launder(layer);  // handed back so that the caller need not keep it across the call (it would live in the stack frame)
}

__device__ __noinline__ void rope_table(float pos, const float* inv_freq, float* rope) {
  // cos/sin are evaluated in fp32 from pos * inv_freq and rounded to bf16 before use (lm_ops.cu)
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int i = lane; i < 64; i += 32) {
    float sn, cs;
    sincosf(pos * inv_freq[i], &sn, &cs);
    rope[i] = bf16_round(cs);
    rope[64 + i] = bf16_round(sn);
  }
}

// ------------------------------------------------------------------ the kernel
template <int kB>
__global__ void __launch_bounds__(kThreads, 1) decode_mega_kernel(const MegaParams p_in) {
  // The (non-inlined) phase routines read the parameter block from one shared copy per CTA (LDS); a reference to the
  // kernel argument would make the compiler keep a per-thread copy in LOCAL memory.
  if (threadIdx.x == 0) s_params = p_in;
  __syncthreads();
  const MegaParams& p = p_in;  // the kernel body itself reads the constant bank
  constexpr int kStages = ring_stages(kB);
  uint8_t* ring = mega_smem + ring_offset(kB);
  bf16* act = sm_act();
  uint64_t* full = sm_full();
  uint64_t* empty = sm_empty();
  float* rope = sm_rope();
  int* sent_wo = sm_sent(0);
  int* sent_wd = sm_sent(1);
  int* sent_gu = sm_sent(2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kCW);
    }
    fence_barrier_init();
  }
  if (tid < 32) s_prof[tid] = 0;
  if (tid < 64) reinterpret_cast<uint32_t*>(act + kB * kActPitch)[tid] = 0u;
  const int NL = p.num_layers;
  const int n_attn = kB * kHkv * p.nsplit;   // CTAs 0 .. n_attn-1 also own one attention unit each
  const int n_gemv = (int)gridDim.x;         // every CTA streams its share of the projections (an SM sustains ~50 GB/s
                                             // of bulk copies, so HBM is only saturated with all of them streaming)
  const uint32_t tag0 = *p.tag_base;         // written by the previous launch (stream order)
  if (tid == 0) {
    s_launch_key = tag0 + 1u;
    g_mega_err = p.err_flag;
  }

  // this CTA's shares (4 per-layer matrices + the LM heads): shared, because the producer indexes them dynamically and a
  // dynamically indexed local array would live in the stack frame
  if (tid < 4) s_sl[tid] = slice_rows(mat_rows(tid), mat_unit(tid), mat_rot(tid), cta, n_gemv);
  if (tid == 4) s_sl[4] = slice_rows(p.vpad, 2, 0, cta, n_gemv);
  if (tid == 5) {
    int c = 0, t = 0;
    for (int m = 0; m < 4; ++m) {
      const Slice sm = slice_rows(mat_rows(m), mat_unit(m), mat_rot(m), cta, n_gemv);
      const int tiles = (sm.r1 - sm.r0 + kTileRows - 1) / kTileRows;
      s_cpre[m] = c;
      s_tpre[m] = t;
      c += tiles * ((m == 3 ? kI : kH) / kChunkK);
      t += tiles;
    }
    s_cpre[4] = c;
    s_tpre[4] = t;
  }
  for (int c = tid; c < n_gemv; c += kThreads) {
    sent_wo[c] = last_tile_row(slice_rows(mat_rows(1), mat_unit(1), mat_rot(1), c, n_gemv)) >> 1;
    sent_wd[c] = last_tile_row(slice_rows(mat_rows(3), mat_unit(3), mat_rot(3), c, n_gemv)) >> 1;
    sent_gu[c] = last_tile_row(slice_rows(mat_rows(2), mat_unit(2), mat_rot(2), c, n_gemv)) >> 2;
  }
  __syncthreads();

  Ring rg{ring, full, empty, 0u};  // (the producer's view; the consumers carry their chunk count in `cstate`)
  const Slice* sl = s_sl;
  const Slice s_heads = s_sl[4];

  if (warp == kCW) {
    // ---------------- producer warp: the CTA's share of every matrix of the step, in consumption order
    const int total = NL * 4 + 1;
#pragma unroll 1
    for (int mi = 0; mi < total; ++mi) {
      const int m = mi & 3;
      const bf16* w;
      Slice s;
      int K = kH;
      if (mi == total - 1) {
        w = p.heads;
        s = s_heads;
      } else {
        const mtts_lm_layer& L = p.layers[mi >> 2];
        w = reinterpret_cast<const bf16*>(m == 0 ? L.wqkv : (m == 1 ? L.wo : (m == 2 ? L.wgu : L.wd)));
        s = sl[m];
        if (m == 3) K = kI;
      }
      const int kch = K / kChunkK;
#pragma unroll 1
      for (int r = s.r0; r < s.r1; r += kTileRows) {
        const int nrows = min(kTileRows, s.r1 - r);
#pragma unroll 1
        for (int kc = 0; kc < kch; ++kc) {
          const uint32_t st = rg.it % kStages, round = rg.it / kStages;
          if (round > 0) mbar_wait(&rg.empty[st], (round - 1) & 1);
          if (lane == 0) mbar_arrive_expect_tx(&rg.full[st], (uint32_t)nrows * kRowBytes);
          __syncwarp();
          if (lane < nrows)
            bulk_g2s(rg.stages + st * kStageBytes + lane * kRowPitch, w + (size_t)(r + lane) * K + (size_t)kc * kChunkK,
                     kRowBytes, &rg.full[st], kEvictFirst);  // evict-normal instead: 13 % slower
          ++rg.it;
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
#ifndef MTTS_MEGA_UNIT_STRIDE
#define MTTS_MEGA_UNIT_STRIDE 1
#endif
  const bool has_unit = (cta % MTTS_MEGA_UNIT_STRIDE) == 0 && cta / MTTS_MEGA_UNIT_STRIDE < n_attn;
  if (tid == 0) s_unit = has_unit ? cta / MTTS_MEGA_UNIT_STRIDE : -1;
  if (has_unit) {
    const int unit_id = cta / MTTS_MEGA_UNIT_STRIDE;
    const int b = unit_id / (p.nsplit * kHkv);
    const int pos = p.positions[b];
    if (tid == 0) {
      AttnUnit un;
      const int split = unit_id % p.nsplit;
      un.hk = (unit_id / p.nsplit) % kHkv;
      un.b = b;
      un.pos = pos;
      const int kv_len = pos + 1;
      const int per = ((kv_len + p.nsplit - 1) / p.nsplit + 31) / 32 * 32;
      un.k_begin = min(kv_len, split * per);
      const int k_end = min(kv_len, un.k_begin + per);
      un.owner = (pos >= un.k_begin && pos < k_end) ? 1 : 0;
      un.k_end = min(k_end, pos);
      s_un = un;
    }
    if (warp == 0) rope_table((float)pos, p.inv_freq, rope);
  }
  // (s_unit, s_un, s_al and rope are first read behind the consumer barriers of the first phase)
#ifdef MTTS_MEGA_PROFILE
  const bool prof = p.prof != nullptr && cta == 0 && tid == 0;
  long long tprev = prof ? clock64() : 0;
#endif
#ifdef MTTS_MEGA_TRACE
#define MEGA_TRACE(slot)                                                            \
  if (p.prof != nullptr && tid == 0 && l == 5) {                                    \
    unsigned long long gt;                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));                          \
    p.prof[32 + cta * 16 + slot] = (long long)gt;                                   \
  }
#else
#define MEGA_TRACE(slot)
#endif
#ifdef MTTS_MEGA_PROFILE
#define MEGA_TICK(slot)                       \
  if (prof) {                                 \
    const long long tn = clock64();           \
    s_prof[slot] += tn - tprev;               \
    tprev = tn;                               \
  }                                           \
  MEGA_TRACE(slot)
#else
#define MEGA_TICK(slot)
#endif
  // Nothing but the layer index lives across the phase calls.
#pragma unroll 1
  for (int l = 0; l < p.num_layers; ++l) {
    MEGA_TICK(13)
    if (threadIdx.x == 0 && s_unit >= 0) {
      const mtts_lm_layer& L = p.layers[l];
      s_al = AttnLayer{reinterpret_cast<const bf16*>(L.k_pool), reinterpret_cast<const bf16*>(L.v_pool),
                       reinterpret_cast<const bf16*>(L.q_norm), reinterpret_cast<const bf16*>(L.k_norm)};
    }
    l = stage_norm<kB>(0, l);
    MEGA_TICK(0)
    l = consume_matrix<kB>(l, EPI_QKV);
    MEGA_TICK(1)
    if (s_unit >= 0) l = attention_layer(l);
    MEGA_TICK(3)
    l = stage_attn_out<kB>(l);
    MEGA_TICK(5)
    l = consume_matrix<kB>(l, EPI_WO);
    MEGA_TICK(6)
    l = stage_norm<kB>(1, l);
    MEGA_TICK(8)
    l = consume_matrix<kB>(l, EPI_GU);
    MEGA_TICK(9)
    l = stage_h<kB>(l);
    MEGA_TICK(11)
    l = consume_matrix<kB>(l, EPI_WD);
    MEGA_TICK(12)
  }
  MEGA_TICK(13)
  stage_norm<kB>(2, p.num_layers);  // final norm + LM heads
  MEGA_TICK(14)
  consume_matrix<kB>(p.num_layers, EPI_HEADS);
  MEGA_TICK(15)
#undef MEGA_TICK
#undef MEGA_TRACE
  // CTA 0 can only be here after it consumed words of every CTA, i.e. after every CTA has read tag_base
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *p.tag_base = s_launch_key - 1u + (uint32_t)(p.num_layers + 1) * kTagsPerLayer;
#ifdef MTTS_MEGA_PROFILE
    if (p.prof)
      for (int i = 0; i < 32; ++i) p.prof[i] += s_prof[i];
#endif
  }
}

typedef void (*MegaKernel)(const MegaParams);
MegaKernel mega_kernel_for(int B) {
  switch (B) {
    case 1: return decode_mega_kernel<1>;
    case 2: return decode_mega_kernel<2>;
    case 3: return decode_mega_kernel<3>;
    default: return decode_mega_kernel<4>;
  }
}

struct MegaLayout { size_t x0, x1, qkv, h, ws, tag, total; };
MegaLayout mega_layout(int B, int nsplit) {
  MegaLayout m;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  m.tag = take(256);
  m.x0 = take((size_t)B * (kH / 2) * 8);
  m.x1 = take((size_t)B * (kH / 2) * 8);
  m.qkv = take((size_t)B * (kNQKV / 2) * 8);
  m.h = take((size_t)B * (kI / 2) * 8);
  m.ws = take((size_t)B * kHkv * nsplit * kG * kWsStride * 8);
  m.total = o;
  return m;
}

}  // namespace

int mtts_configure_decode_mega() {
  for (int B = 1; B <= kMaxB; ++B) {
    cudaError_t e = cudaFuncSetAttribute(mega_kernel_for(B), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total(B));
    if (e != cudaSuccess) return mtts_set_error(MTTS_ERR_CUDA, "decode_mega: smem attribute: %s", cudaGetErrorString(e));
  }
  return MTTS_OK;
}

extern "C" long long mtts_decode_mega_workspace_bytes(int B, int nsplit) {
  if (B <= 0 || nsplit <= 0) return 0;
  return (long long)mega_layout(B, nsplit).total;
}

extern "C" int mtts_decode_mega_supported(int hidden, int intermediate, int q_heads, int kv_heads, int head_dim, int B) {
  return hidden == kH && intermediate == kI && q_heads == kHq && kv_heads == kHkv && head_dim == kD && B >= 1 && B <= kMaxB;
}

extern "C" int mtts_decode_mega(const mtts_decode_mega_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(a != nullptr, "mtts_decode_mega: null args");
  MTTS_REQUIRE(mtts_decode_mega_supported(a->hidden, a->intermediate, a->num_q_heads, a->num_kv_heads, a->head_dim, a->B),
               "mtts_decode_mega: unsupported shape (hidden %d, intermediate %d, heads %d/%d x %d, batch %d); "
               "supported: 2048/6144, 16/8 x 128, batch 1..4",
               a->hidden, a->intermediate, a->num_q_heads, a->num_kv_heads, a->head_dim, a->B);
  MTTS_REQUIRE(a->layers && a->heads && a->final_norm && a->inv_freq && a->positions && a->x && a->logits && a->workspace,
               "mtts_decode_mega: null pointer");
  MTTS_REQUIRE(a->num_layers >= 1 && a->vpad > 0 && a->vpad % 2 == 0, "mtts_decode_mega: bad num_layers / vpad");
  MTTS_REQUIRE(a->page_size > 0 && (a->page_size & (a->page_size - 1)) == 0, "mtts_decode_mega: page_size must be a power of two");
  MTTS_REQUIRE(a->nsplit >= 1 && a->nsplit <= 16, "mtts_decode_mega: nsplit must be in [1,16]");
  MTTS_REQUIRE(a->B * a->num_kv_heads * a->nsplit <= mtts_num_sms(),
               "mtts_decode_mega: B * kv_heads * nsplit attention units exceed the SM count (one unit per CTA)");
  MTTS_REQUIRE(a->workspace_bytes >= mtts_decode_mega_workspace_bytes(a->B, a->nsplit), "mtts_decode_mega: workspace too small");
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "mtts_decode_mega: workspace must be 256-byte aligned");
  MTTS_REQUIRE(a->ld_logits >= a->vpad, "mtts_decode_mega: ld_logits < vpad");
  const MegaLayout lay = mega_layout(a->B, a->nsplit);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  MegaParams p;
  p.layers = a->layers; p.num_layers = a->num_layers;
  p.heads = reinterpret_cast<const bf16*>(a->heads); p.vpad = a->vpad;
  p.final_norm = reinterpret_cast<const bf16*>(a->final_norm);
  p.inv_freq = a->inv_freq; p.positions = a->positions; p.block_table = a->block_table;
  p.max_pages = a->max_pages; p.num_pages = a->num_pages;
  int shift = 0;
  while ((1 << shift) < a->page_size) ++shift;
  p.page_shift = shift;
  p.x_in = reinterpret_cast<const bf16*>(a->x);
  p.x_ll[0] = reinterpret_cast<uint2*>(ws + lay.x0);
  p.x_ll[1] = reinterpret_cast<uint2*>(ws + lay.x1);
  p.qkv_ll = reinterpret_cast<uint2*>(ws + lay.qkv);
  p.h_ll = reinterpret_cast<uint2*>(ws + lay.h);
  p.ws_ll = reinterpret_cast<uint2*>(ws + lay.ws);
  p.tag_base = reinterpret_cast<unsigned int*>(ws + lay.tag);
  p.logits = reinterpret_cast<bf16*>(a->logits); p.ld_logits = a->ld_logits;
  p.nsplit = a->nsplit; p.eps = a->eps;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)kD);
  p.err_flag = a->err_flag;
  p.prof = a->profile_cycles;
  {
    // one sentinel warp per CTA before the CTA-wide verified load (MTTS_MEGA_SENTINEL=0: every thread polls the words
    // it needs directly — 3.5 % slower at batch 1 and 4, 1 % at batch 2)
    static int sen = -1;
    if (sen < 0) {
      const char* e = getenv("MTTS_MEGA_SENTINEL");
      sen = (e && e[0] == '0') ? 0 : 1;
    }
    p.sentinel = sen;
  }

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)mtts_num_sms());
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)smem_total(a->B);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch fails (never a hung poll loop)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MTTS_REQUIRE(mtts_num_sms() <= kMaxCtas, "mtts_decode_mega: more SMs than the sentinel tables hold");
  MTTS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mega_kernel_for(a->B), p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
