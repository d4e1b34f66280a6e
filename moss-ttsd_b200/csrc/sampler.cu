// Delay-pattern sampler: per-channel logit masks + HF logits processors + draw, and the per-row state
// machine of CustomMixin._sample, all on the device (the reference needs ~60 launches and two host syncs
// per step for this, modeling_asteroid.py:123-169).
//
//   sample_scan/finish one CTA per (row, channel, 4096-logit slice): masks (modeling_asteroid.py:124-128), repetition penalty ->
//                      temperature -> top-k -> top-p in HF's order and semantics (:95-106,129; HF
//                      RepetitionPenaltyLogitsProcessor / TemperatureLogitsWarper / TopKLogitsWarper /
//                      TopPLogitsWarper), then multinomial draw or argmax (:131-138).
//   delay_step_kernel  wind-down trigger, teacher forcing of the delayed prompt tail, wind-down fill,
//                      finished-row fill, append, counters and stopping (:140-169).
//
// Repetition penalty needs "was token j ever in this channel's history" — a per-(row, channel) bitmap that
// the step kernel updates, instead of a gather/scatter over a history that grows to 16k entries.
// Top-k never sorts the vocabulary: the k-th largest of the 512 per-thread maxima is a lower bound of the
// k-th largest score, so one more pass collects the (few) candidates above it and only those are sorted.
#include "common.cuh"
#include "mtts_internal.h"
#include <string.h>

namespace {

constexpr int kThreads = 512;
constexpr int kCap = 2048;  // candidate list capacity

// ---- Philox4x32-10 (counter-based; one independent stream per (step, row, channel))
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t step, uint32_t stream_id) {
  uint32_t c[4] = {step, stream_id, 0x6d747473u, 0u};
  uint32_t k[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  return (c[0] >> 8) * (1.0f / 16777216.0f);  // [0, 1)
}

struct ScoreCtx {
  const bf16* lg;
  const uint32_t* seen;
  int mask_idx;
  bool has_rep, has_temp;
  float pen, temp;
};
__device__ __forceinline__ float score_at(const ScoreCtx& c, int j) {
  float s = __bfloat162float(c.lg[j]);  // `.clone().float()` of the bf16 head output (:123)
  if (j == c.mask_idx) s = -INFINITY;
  if (c.has_rep && ((c.seen[j >> 5] >> (j & 31)) & 1u)) s = s < 0.f ? s * c.pen : s / c.pen;
  if (c.has_temp) s = s / c.temp;
  return s;
}

__device__ __forceinline__ bool better(float v, int i, float ov, int oi) { return v > ov || (v == ov && i < oi); }
// Candidate order for top-p: descending value, and among EQUAL values descending index. bf16 logits tie often; HF
// sorts ascending (stable), so within a tie group the lower index has the smaller cumulative mass and is removed
// first — i.e. in descending order the higher index must come first.
__device__ __forceinline__ bool sorts_before(float v, int i, float ov, int oi) { return v > ov || (v == ov && i > oi); }

// bitonic sort of (val, idx) in shared memory, descending by val then ascending idx; n is a power of two
__device__ void bitonic_sort_desc(float* val, int* idx, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int x = t ^ j;
        if (x > t) {
          const bool up = (t & k) == 0;  // "up" = this run sorted in our target (descending) order
          const bool t_first = sorts_before(val[t], idx[t], val[x], idx[x]);
          if (up ? !t_first : t_first) {
            const float tv = val[t]; val[t] = val[x]; val[x] = tv;
            const int ti = idx[t]; idx[t] = idx[x]; idx[x] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The 152,697-wide channel-0 row is cut into slices of kSlice logits, one CTA each, so that a batch-1 step is not a
// single CTA crawling over 300 KB (that cost 180 us); small channels are one slice. Two phases, each finished by
// the last CTA of a (row, channel) to arrive (atomic ticket, no spinning):
//   scan    every slice computes its per-thread maxima. Greedy channels: slice argmax -> last arriver reduces -> token.
//           Sampled channels: each slice reports its kReport largest thread maxima; the k-th largest of all reported
//           values is a lower bound of the k-th largest score -> threshold.
//   finish  (sampled channels only) every slice pushes its scores >= threshold into a small global candidate list;
//           the last arriver sorts it and applies top-k (ties kept) / top-p / the Philox draw.
// ------------------------------------------------------------------------------------------------
constexpr int kSlice = kThreads * 8;  // 4096 logits per CTA, 16-byte loads
constexpr int kReport = 256;  // reported scores per slice: the 16 largest thread maxima of each of its 16 warps
constexpr int kMaxSlices = 64;

struct SampleWs {
  float* slice_val;   // [B][C][kMaxSlices]
  int* slice_idx;     // [B][C][kMaxSlices]
  float* reported;    // [B][kMaxSlices (all channels' slices, launch order)][kReport]
  float* thr;         // [B][C][4]: threshold, global max, global sum of exp(score - max), unused
  int* tickets;       // [B][C][2]  (scan, finish) zero between launches
  int* cand_count;    // [B][C]     zero between launches
  int* redo;          // [B][C]     set by the finish phase when a nucleus does not fit the candidate list; cleared by sample_exact
  float* cand_val;    // [B][C][kCap]
  int* cand_idx;      // [B][C][kCap]
};

struct SampleParams2 {
  const bf16* logits;
  long long ld;
  mtts_sampler_config cfg;
  const uint32_t* seen;
  const int* step_ptr;
  const int* row_ctl;  // optional [B][4] = {step0, P, max_length, eos_at}: per-row step origin (continuous batching)
  const unsigned long long* seed_ptr;  // device-resident so that a captured graph can be re-used with a new seed
  long long* out_tokens;
  int* err_flag;
  SampleWs ws;
  int total_slices;
  unsigned int exact_mask;  // channels sampled by sample_exact_kernel alone (wide top-k / no filter on a wide vocabulary)
  unsigned char exact_channels[8];  // the channels sample_exact_kernel is launched for (vocab > kCap, sampled)
  int chunk;  // kSlice-wide pieces one CTA scans (> 1 only when every channel is greedy and the batch is large)
  unsigned char slice_channel[kMaxSlices];
  unsigned char slice_index[kMaxSlices];
  unsigned char slices_of[8];
};

__device__ __forceinline__ ScoreCtx make_ctx(const SampleParams2& p, int b, int c, int step) {
  const mtts_sampler_config& cfg = p.cfg;
  ScoreCtx sc;
  sc.lg = p.logits + (long long)b * p.ld + cfg.logit_offset[c];
  sc.seen = p.seen + (long long)b * cfg.seen_words_per_row + cfg.seen_offset_words[c];
  sc.mask_idx = -1;
  if (c != 0 && step >= c) sc.mask_idx = cfg.pad_token;                        // channel c is live: pad is illegal
  if (c == 0 && step <= cfg.channels - 2) sc.mask_idx = cfg.eos_mask_token;    // no EOS while the prompt tail is forced
  sc.has_rep = cfg.has_rep[c] != 0;
  sc.pen = cfg.rep_penalty[c];
  sc.has_temp = cfg.has_temp[c] != 0;
  sc.temp = cfg.temperature[c];
  return sc;
}

// processed scores of logits j0 .. j0+7 (j0 % 8 == 0); entries at or beyond V come back as -inf
__device__ __forceinline__ void scores8(const ScoreCtx& c, int j0, int V, float (&s)[8]) {
  if (j0 >= V) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = -INFINITY;
    return;
  }
  if (j0 + 8 <= V) {
    const uint4 u = *reinterpret_cast<const uint4*>(c.lg + j0);
    s[0] = bf16lo(u.x); s[1] = bf16hi(u.x); s[2] = bf16lo(u.y); s[3] = bf16hi(u.y);
    s[4] = bf16lo(u.z); s[5] = bf16hi(u.z); s[6] = bf16lo(u.w); s[7] = bf16hi(u.w);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = (j0 + e < V) ? __bfloat162float(c.lg[j0 + e]) : -INFINITY;
  }
  const uint32_t bits = c.has_rep ? (c.seen[j0 >> 5] >> (j0 & 31)) & 0xffu : 0u;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = s[e];
    if (j0 + e == c.mask_idx) v = -INFINITY;
    if ((bits >> e) & 1u) v = v < 0.f ? v * c.pen : v / c.pen;
    if (c.has_temp) v = v / c.temp;
    s[e] = v;
  }
}

__device__ __forceinline__ void block_argmax(float& bv, int& bi, float* s_val, int* s_idx) {
  const int tid = threadIdx.x;
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
  __syncthreads();
  if (tid < 32) {
    bv = tid < kThreads / 32 ? s_val[tid] : -INFINITY;
    bi = tid < kThreads / 32 ? s_idx[tid] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  }
}

// (two CTAs per SM: at 80 registers and the default shared-memory carve-out the 11 520 slice CTAs of a batch-256 step ran one
// per SM — 78 waves; ncu: launch__occupancy_limit_registers = launch__occupancy_limit_shared_mem = 1)
__global__ void __launch_bounds__(kThreads, 2) sample_scan_kernel(const SampleParams2 p) {
  __shared__ float s_val[kThreads];
  __shared__ int s_idx[kThreads];
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, c = p.slice_channel[blockIdx.y], slice = p.slice_index[blockIdx.y];
  const mtts_sampler_config& cfg = p.cfg;
  if ((p.exact_mask >> c) & 1u) return;  // sampled by sample_exact_kernel
  const int V = cfg.vocab[c], S = p.slices_of[c];
  const int tid = threadIdx.x;
  const int step = *p.step_ptr - (p.row_ctl ? p.row_ctl[b * 4] : 0);
  const ScoreCtx sc = make_ctx(p, b, c, step);
  const int j0 = slice * p.chunk * kSlice + tid * 8;
  const long long bc = (long long)b * cfg.channels + c;
  const bool greedy = !cfg.do_sample[c];
  float sv[8];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if (!greedy) {
    for (int q = 0; q < p.chunk; ++q) {  // `chunk` pieces of 4096 logits per CTA at large batch (fewer, fatter CTAs)
      const int jq = j0 + q * kSlice;
      if (jq >= V) break;
      scores8(sc, jq, V, sv);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (jq + e < V && better(sv[e], jq + e, bv, bi)) { bv = sv[e]; bi = jq + e; }
    }
  } else {
    // Greedy: the scan is issue-bound (ncu: 386 instructions per warp for 8 logits per thread), so a group of 8 raw
    // logits is first reduced with plain max and only searched for its index when it beats the running best; groups
    // that contain a masked or penalised entry (or any group under a temperature) take the general path. Large
    // batches walk `chunk` pieces per CTA, four 16-byte loads in flight per thread.
    for (int it = 0; it < p.chunk; it += 4) {
      uint4 u[4];
      bool fast[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jq = j0 + (it + q) * kSlice;
        fast[q] = it + q < p.chunk && jq + 8 <= V;
        if (fast[q]) u[q] = *reinterpret_cast<const uint4*>(sc.lg + jq);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jq = j0 + (it + q) * kSlice;
        if (it + q >= p.chunk || jq >= V) continue;
        bool plain = fast[q] && !sc.has_temp && !(sc.mask_idx >= jq && sc.mask_idx < jq + 8);
        if (plain && sc.has_rep) plain = ((sc.seen[jq >> 5] >> (jq & 31)) & 0xffu) == 0u;
        if (plain) {
          const float v0 = bf16lo(u[q].x), v1 = bf16hi(u[q].x), v2 = bf16lo(u[q].y), v3 = bf16hi(u[q].y);
          const float v4 = bf16lo(u[q].z), v5 = bf16hi(u[q].z), v6 = bf16lo(u[q].w), v7 = bf16hi(u[q].w);
          const float m = fmaxf(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)), fmaxf(fmaxf(v4, v5), fmaxf(v6, v7)));
          if (m > bv) {  // pieces are visited in ascending index order: an equal maximum further on never wins
            bv = m;
            bi = jq + (v0 == m ? 0 : v1 == m ? 1 : v2 == m ? 2 : v3 == m ? 3 : v4 == m ? 4 : v5 == m ? 5 : v6 == m ? 6 : 7);
          }
        } else {
          scores8(sc, jq, V, sv);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (jq + e < V && better(sv[e], jq + e, bv, bi)) { bv = sv[e]; bi = jq + e; }
        }
      }
    }
  }
  float wmax = -INFINITY;  // (sampled channels) largest score of this warp's 256 logits
  if (greedy) {
    block_argmax(bv, bi, s_val, s_idx);
    if (tid == 0) {
      p.ws.slice_val[bc * kMaxSlices + slice] = bv;
      p.ws.slice_idx[bc * kMaxSlices + slice] = bi;
    }
  } else {
    // Reports of a slice: the thread maxima are sorted inside every warp (bitonic network over shuffles: no barrier) and
    // each warp writes its 16 largest — kReport = 256 scores per slice, no block-wide sort (the first version's 512-entry
    // bitonic sort, 45 barrier stages per slice on 9728 slices at batch 256, made the scan 984 us of a sampled step).
    // Any set of real scores gives a valid threshold (the k-th largest of a SUBSET is a lower bound of the k-th largest
    // score); it has to stay tight when the whole top k sits in a few warps — a TTS step puts all of channel 0's mass
    // into the 1024 contiguous speech tokens, four warps of one slice (with only 4 reports per warp the 50th largest
    // report was a cold logit and every score of the row became a candidate: device flag 3).
    const int lane = tid & 31, wid = tid >> 5;
    float v = bv;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, j);
        const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);  // descending over the whole warp at k = 32
        v = keep_max ? fmaxf(v, o) : fminf(v, o);
      }
    if (lane < 16) p.ws.reported[((long long)b * kMaxSlices + blockIdx.y) * kReport + wid * 16 + lane] = v;
    wmax = __shfl_sync(0xffffffffu, v, 0);
  }
  const bool need_mass = !greedy && S > 1 && cfg.top_k[c] <= 0;  // only a nucleus over the WHOLE vocabulary needs the row's softmax mass
  if (need_mass) {
    // slice softmax statistics (needed for top-p over the full vocabulary when no top-k precedes it)
    if ((tid & 31) == 0) s_val[tid >> 5] = wmax;
    __syncthreads();
    float smax = s_val[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) smax = fmaxf(smax, s_val[w]);
    float se = 0.f;
    for (int q = 0; q < p.chunk; ++q) {
      const int jq = j0 + q * kSlice;
      if (jq >= V) break;
      scores8(sc, jq, V, sv);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (jq + e < V && sv[e] > -INFINITY) se += expf(sv[e] - smax);
    }
    __shared__ float s_red[33];
    se = block_sum(se, s_red);
    if (tid == 0) {
      p.ws.slice_val[bc * kMaxSlices + slice] = smax;
      p.ws.slice_idx[bc * kMaxSlices + slice] = __float_as_int(se);
    }
  }
  // ---- ticket: the last slice of this (row, channel) finishes the phase
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(p.ws.tickets + bc * 2, 1);
    s_last = (prev == S - 1);
    if (s_last) p.ws.tickets[bc * 2] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (greedy) {
    bv = -INFINITY;
    bi = 0x7fffffff;
    if (tid < S) {
      bv = __ldcg(p.ws.slice_val + bc * kMaxSlices + tid);
      bi = __ldcg(p.ws.slice_idx + bc * kMaxSlices + tid);
    }
    __syncthreads();
    block_argmax(bv, bi, s_val, s_idx);
    if (tid == 0) p.out_tokens[bc] = bi == 0x7fffffff ? 0 : bi;  // torch.argmax: lowest index on ties
    return;
  }
  // sampled channel: threshold = k-th largest reported maximum (a lower bound of the k-th largest score). Without
  // top-k (top-p over the whole vocabulary) the candidate list is the ~1500 largest scores; the nucleus is cut inside
  // it using the GLOBAL softmax mass (flagged if the nucleus does not fit).
  float thr = -INFINITY;
  int k = cfg.top_k[c] > 0 ? min(cfg.top_k[c], V) : V;
  if (need_mass) {
    float gm = -INFINITY;
    for (int t = 0; t < S; ++t) gm = fmaxf(gm, __ldcg(p.ws.slice_val + bc * kMaxSlices + t));
    float gz = 0.f;
    for (int t = 0; t < S; ++t)
      gz += __int_as_float(__ldcg(p.ws.slice_idx + bc * kMaxSlices + t)) * expf(__ldcg(p.ws.slice_val + bc * kMaxSlices + t) - gm);
    if (tid == 0) { p.ws.thr[bc * 4 + 1] = gm; p.ws.thr[bc * 4 + 2] = gz; }
    if (cfg.top_k[c] <= 0) k = min(S * kReport, (kCap * 3) / 4);
  }
  // k-th largest of the S * kReport reported scores, exactly, by a most-significant-bit-first radix select on
  // order-preserving integer keys (32 rounds of count-and-decide over values held in registers; no sort, no big buffer)
  const int n_rep = S * kReport;
  if (k <= n_rep) {
    const float* rep = p.ws.reported + ((long long)b * kMaxSlices + (blockIdx.y - slice)) * kReport;  // this channel's slices are contiguous
    constexpr int kPer = kMaxSlices * kReport / kThreads;  // 32
    uint32_t key[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int t = tid + i * kThreads;
      uint32_t u = 0u;  // below every real key (-inf maps to 0x007fffff)
      if (t < n_rep) {
        u = __float_as_uint(__ldcg(rep + t));
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      }
      key[i] = u;
    }
    __shared__ int s_cnt[2];
    uint32_t prefix = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      if (tid == 0) s_cnt[bit & 1] = 0;
      __syncthreads();
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < kPer; ++i) cnt += key[i] >= cand ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if ((tid & 31) == 0 && cnt) atomicAdd(&s_cnt[bit & 1], cnt);
      __syncthreads();
      if (s_cnt[bit & 1] >= k) prefix = cand;
    }
    const uint32_t u = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
    thr = __uint_as_float(u);
  }
  if (tid == 0) p.ws.thr[bc * 4] = thr;
}

__global__ void __launch_bounds__(kThreads) sample_finish_kernel(const SampleParams2 p) {
  __shared__ float s_val[kCap];
  __shared__ int s_idx[kCap];
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, c = p.slice_channel[blockIdx.y], slice = p.slice_index[blockIdx.y];
  const mtts_sampler_config& cfg = p.cfg;
  if (!cfg.do_sample[c] || ((p.exact_mask >> c) & 1u)) return;
  const int V = cfg.vocab[c], S = p.slices_of[c];
  const int tid = threadIdx.x;
  const int step = *p.step_ptr - (p.row_ctl ? p.row_ctl[b * 4] : 0);
  const ScoreCtx sc = make_ctx(p, b, c, step);
  const long long bc = (long long)b * cfg.channels + c;
  const float thr = __ldcg(p.ws.thr + bc * 4);
  const int j0 = slice * p.chunk * kSlice + tid * 8;
  float sv[8];
  scores8(sc, j0, V, sv);
  int n, npad = 1;
  if (S == 1 && V <= kCap) {
    // a one-slice channel (the 1025-way speech channels): this CTA holds the whole row — the scores go straight into the
    // sort buffer (the candidate list in global memory cost one atomic per score on a single counter: 1025 serialised
    // atomics per (row, channel))
    // ... and only the scores at or above the scan's threshold (a lower bound of the k-th largest) are kept: ~50-150 of 1025,
    // a 128/256-entry sort instead of a 2048-entry one (ncu: the finish kernel was issue-bound, 310 M instructions per step)
    __shared__ int s_n;
    if (tid == 0) s_n = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (j0 + e < V && sv[e] >= thr) {
        const int pos = atomicAdd(&s_n, 1);
        s_val[pos] = sv[e];
        s_idx[pos] = j0 + e;
      }
    __syncthreads();
    n = s_n;
    while (npad < n) npad <<= 1;
    for (int t = n + tid; t < npad; t += kThreads) { s_val[t] = -INFINITY; s_idx[t] = -1; }
    __syncthreads();
  } else {
    for (int q = 0; q < p.chunk; ++q) {
      const int jq = j0 + q * kSlice;
      if (jq >= V) break;
      if (q > 0) scores8(sc, jq, V, sv);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (jq + e < V && sv[e] >= thr) {
          const int pos = atomicAdd(p.ws.cand_count + bc, 1);
          if (pos < kCap) {
            __stcg(p.ws.cand_val + bc * kCap + pos, sv[e]);
            __stcg(p.ws.cand_idx + bc * kCap + pos, jq + e);
          }
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(p.ws.tickets + bc * 2 + 1, 1);
      s_last = (prev == S - 1);
      if (s_last) p.ws.tickets[bc * 2 + 1] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    n = __ldcg(p.ws.cand_count + bc);
    __syncthreads();
    if (tid == 0) p.ws.cand_count[bc] = 0;
    if (n > kCap) {
      if (tid == 0 && p.err_flag) *p.err_flag = 3;  // candidate overflow (pathological ties); truncated
      n = kCap;
    }
    while (npad < n) npad <<= 1;
    for (int t = tid; t < npad; t += kThreads) {
      s_val[t] = t < n ? __ldcg(p.ws.cand_val + bc * kCap + t) : -INFINITY;
      s_idx[t] = t < n ? __ldcg(p.ws.cand_idx + bc * kCap + t) : -1;
    }
    __syncthreads();
  }
  bitonic_sort_desc(s_val, s_idx, npad);
  const int k = cfg.top_k[c] > 0 ? min(cfg.top_k[c], V) : V;

  // ---------------- top-k (ties with the k-th value are kept, HF: scores < kth -> -inf), top-p, draw
  if (tid < 32) {
    const int lane = tid;
    int nk = n;
    if (cfg.top_k[c] > 0 && k < n) {
      const float kth = s_val[k - 1];
      int cnt = 0;
      for (int t = lane; t < n; t += 32) cnt += (s_val[t] >= kth) ? 1 : 0;
      cnt = (int)warp_sum((float)cnt);
      nk = cnt;
    }
    // -inf scores have probability 0 and are never drawn: drop them (but keep at least one entry)
    {
      int cnt = 0;
      for (int t = lane; t < nk; t += 32) cnt += (s_val[t] > -INFINITY) ? 1 : 0;
      cnt = (int)warp_sum((float)cnt);
      nk = max(cnt, 1);
    }
    const float vmax = s_val[0];
    __syncwarp();  // every lane has read the maximum before lane 0 overwrites it below
    float part = 0.f;
    const int chunk = (nk + 31) / 32;
    const int t0 = lane * chunk, t1 = min(nk, t0 + chunk);
    for (int t = t0; t < t1; ++t) {
      const float e = expf(s_val[t] - vmax);
      s_val[t] = e;
      part += e;
    }
    float incl = part;
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    float Z = __shfl_sync(0xffffffffu, incl, 31);
    const bool global_mass = (cfg.top_k[c] <= 0 && S > 1);
    if (global_mass) Z = __ldcg(p.ws.thr + bc * 4 + 2);  // softmax mass of the WHOLE row (relative to the same maximum)
    int keep = nk;
    if (cfg.has_top_p[c]) {
      // HF: sort ascending, cum = cumsum(softmax); remove cum <= 1 - top_p; always keep the largest.
      // In descending order: position t is removed iff (mass of t and everything after it) / Z <= 1 - top_p.
      const float limit = (1.0f - cfg.top_p[c]);
      float before = incl - part;
      int my_keep = 0;
      for (int t = t0; t < t1; ++t) {
        const float tail = (Z - before) / Z;
        if (t == 0 || tail > limit) my_keep = t + 1;
        before += s_val[t];
      }
      for (int o = 16; o > 0; o >>= 1) my_keep = max(my_keep, __shfl_xor_sync(0xffffffffu, my_keep, o));
      keep = max(my_keep, 1);
      // full-vocabulary nucleus larger than the candidate list: this (row, channel) is drawn again by sample_exact_kernel
      if (global_mass && keep >= nk && n >= (kCap * 3) / 4 && lane == 0) p.ws.redo[bc] = 1;
    }
    float kpart = 0.f;
    for (int t = t0; t < min(t1, keep); ++t) kpart += s_val[t];
    float kincl = kpart;
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, kincl, o);
      if (lane >= o) kincl += up;
    }
    const float Zk = __shfl_sync(0xffffffffu, kincl, 31);
    const float u = philox_uniform(*p.seed_ptr, (uint32_t)step, (uint32_t)(b * 8 + c));
    const float target = u * Zk;
    const float lo = kincl - kpart;
    int choice = -1;
    if (target >= lo && target < kincl) {
      float acc = lo;
      for (int t = t0; t < min(t1, keep); ++t) {
        acc += s_val[t];
        if (target < acc) { choice = t; break; }
      }
      if (choice < 0) choice = min(t1, keep) - 1;
    }
    for (int o = 16; o > 0; o >>= 1) choice = max(choice, __shfl_xor_sync(0xffffffffu, choice, o));
    if (choice < 0) choice = keep - 1;  // rounding at the very end of the CDF
    if (lane == 0) p.out_tokens[bc] = s_idx[choice];
  }
}

// ------------------------------------------------------------------------------------------------
// Exact draw over a wide vocabulary (the 152,697-way text channel) for what the candidate-list path cannot represent:
// top_k above the list's reach, no filter at all (plain temperature sampling), or a nucleus that outgrows the list. One
// CTA per (row, channel) walks the row several times and never sorts or stores it:
//   top-k   the k-th largest score by a radix select on order-preserving integer keys (4 passes, 256-bin counts);
//           ties with the k-th value are kept (HF TopKLogitsWarper: scores < kth -> -inf)
//   top-p   HF removes, in ascending order, every token whose cumulative probability is <= 1 - top_p: a token stays iff
//           the mass of all kept-by-top-k scores <= its own exceeds (1 - top_p) Z -> the smallest such score is found by
//           bisection on the integer keys (32 passes, fixed-order reductions: deterministic); equal scores stay together
//   draw    inverse CDF over the surviving scores, thread-major order (the order only has to be fixed), own Philox stream
// A slow path by design (~40 passes over 300 KB of L2-resident logits per row); the common configurations never get here.
__device__ __forceinline__ uint32_t order_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// fixed-order block sum (warp shuffles, then the 16 warp sums in index order): every thread gets the total
__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(kThreads) sample_exact_kernel(const SampleParams2 p) {
  __shared__ float s_red[kThreads / 32];
  __shared__ int s_hist[256];
  __shared__ uint32_t s_sel[2];
  __shared__ float s_scan[kThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, c = p.exact_channels[blockIdx.y];
  const mtts_sampler_config& cfg = p.cfg;
  const long long bc = (long long)b * cfg.channels + c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!((p.exact_mask >> c) & 1u)) {  // a channel of the candidate-list path: only rows it gave up on
    if (__ldcg(p.ws.redo + bc) == 0) return;
    __syncthreads();
    if (tid == 0) p.ws.redo[bc] = 0;
  }
  const int V = cfg.vocab[c];
  const int step = *p.step_ptr - (p.row_ctl ? p.row_ctl[b * 4] : 0);
  const ScoreCtx sc = make_ctx(p, b, c, step);
  const int pieces = (V + kSlice - 1) / kSlice;
  float sv[8];

  // ---- maximum
  float m = -INFINITY;
  for (int q = 0; q < pieces; ++q) {
    scores8(sc, q * kSlice + tid * 8, V, sv);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, sv[e]);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, s_red[w]);

  // ---- top-k threshold key (0 = keep everything)
  uint32_t kth_key = 0u;
  const int k = cfg.top_k[c] > 0 ? min(cfg.top_k[c], V) : V;
  if (k < V) {
    uint32_t prefix = 0u;
    int want = k;  // rank, among the keys that share `prefix`, of the key we are after (1 = largest)
    for (int shift = 24; shift >= 0; shift -= 8) {
      __syncthreads();
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int q = 0; q < pieces; ++q) {
        const int j0 = q * kSlice + tid * 8;
        scores8(sc, j0, V, sv);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (j0 + e < V) {
            const uint32_t key = order_key(sv[e]);
            if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1);
          }
      }
      __syncthreads();
      if (tid == 0) {
        int acc = 0, d = 255;
        for (; d > 0; --d) {
          if (acc + s_hist[d] >= want) break;
          acc += s_hist[d];
        }
        s_sel[0] = (uint32_t)d;
        s_sel[1] = (uint32_t)(want - acc);
      }
      __syncthreads();
      prefix |= s_sel[0] << shift;
      want = (int)s_sel[1];
    }
    kth_key = prefix;
  }

  // ---- softmax mass of what top-k keeps
  float z = 0.f;
  for (int q = 0; q < pieces; ++q) {
    const int j0 = q * kSlice + tid * 8;
    scores8(sc, j0, V, sv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (j0 + e < V && order_key(sv[e]) >= kth_key) z += expf(sv[e] - m);
  }
  const float Z = block_sum(z, s_red);

  // ---- top-p: the smallest key whose "mass at or below it" exceeds (1 - top_p) Z; the maximum always stays
  uint32_t cut_key = kth_key;
  if (cfg.has_top_p[c] && cfg.top_p[c] < 1.0f) {
    const float limit = (1.0f - cfg.top_p[c]) * Z;
    uint32_t lo = kth_key, hi = order_key(m);  // invariant: the answer is in [lo, hi]; "mass <= hi" = Z > limit, or the maximum is kept anyway
#pragma unroll 1
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      float part = 0.f;
      for (int q = 0; q < pieces; ++q) {
        const int j0 = q * kSlice + tid * 8;
        scores8(sc, j0, V, sv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t key = order_key(sv[e]);
          if (j0 + e < V && key >= kth_key && key <= mid) part += expf(sv[e] - m);
        }
      }
      const float mass = block_sum(part, s_red);
      if (mass > limit) hi = mid; else lo = mid + 1;
    }
    cut_key = lo;
  }

  // ---- draw: inverse CDF over the survivors, thread-major
  float w = 0.f;
  for (int q = 0; q < pieces; ++q) {
    const int j0 = q * kSlice + tid * 8;
    scores8(sc, j0, V, sv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (j0 + e < V && order_key(sv[e]) >= cut_key) w += expf(sv[e] - m);
  }
  float incl = w;
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  __syncthreads();
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  float before = 0.f, total = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) {
    if (i < warp) before += s_scan[i];
    total += s_scan[i];
  }
  const float lo_t = before + incl - w, hi_t = before + incl;
  const float u = philox_uniform(*p.seed_ptr, (uint32_t)step, (uint32_t)(b * 8 + c));
  const float target = u * total;
  if (tid == 0) s_sel[0] = 0xffffffffu;
  __syncthreads();
  // the owner of `target`; rounding at the very end of the CDF falls to the last thread with any mass
  const bool mine = w > 0.f && ((target >= lo_t && target < hi_t) || (target >= hi_t && hi_t == total));
  if (mine) {
    float acc = lo_t;
    int choice = -1, last = -1;
    for (int q = 0; q < pieces && choice < 0; ++q) {
      const int j0 = q * kSlice + tid * 8;
      scores8(sc, j0, V, sv);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (choice < 0 && j0 + e < V && order_key(sv[e]) >= cut_key) {
          acc += expf(sv[e] - m);
          last = j0 + e;
          if (target < acc) choice = j0 + e;
        }
    }
    if (choice < 0) choice = last;
    atomicMin(&s_sel[0], (uint32_t)choice);  // (two owners only if partial sums tie exactly: any of them is a legal draw)
  }
  __syncthreads();
  if (tid == 0) p.out_tokens[bc] = s_sel[0] == 0xffffffffu ? 0 : (long long)s_sel[0];
}

// ------------------------------------------------------------------------------------------------
struct StepParams {
  long long* raw_tokens;     // [B, C] sampled this step (in) / final tokens (out)
  const long long* tf_tail;  // [B, C-1, C] teacher-forced prompt tail = prompt[:, P:P+C-1, :]
  long long* sequences;      // [B, max_len, C] output grid
  long long max_len_rows;    // rows allocated in `sequences`
  int* unfinished;           // [B]
  int* needs_steps;          // [B]
  int* positions;            // [B] next RoPE position / current KV length
  uint32_t* seen;
  int* step_ptr;
  int* unfinished_hist;      // [max_steps] number of unfinished rows after each step
  int hist_len;              // per-row mode: ring length of unfinished_hist
  int* finish_len;           // [B] sequence length (rows) at which the row finished (0 = not yet)
  int B, C;
  const int* dyn;  // device: [0] = prompt rows P, [1] = max_length (so a captured graph survives new prompts)
  // optional [B][4] = {step0, P, max_length, eos_at}: every row runs the state machine from its own step origin with
  // its own prompt length / length limit (continuous batching: a finished row's slot is refilled while the others go
  // on), and `eos_at` > 0 replaces channel 0's token by EOS from sequence row `eos_at` on (a per-request length budget:
  // the row then winds down exactly as if the model had emitted EOS there, modeling_asteroid.py:140-153)
  const int* row_ctl;
  int speech_lo, speech_hi, eos_token, pad_token, has_eos_criteria;
  mtts_sampler_config cfg;
};

__global__ void delay_step_kernel(const StepParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_count;
  const int b = threadIdx.x;
  const int gs = *p.step_ptr;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (b < p.B) {
    const int C = p.C;
    const int* rc = p.row_ctl ? p.row_ctl + b * 4 : nullptr;
    const int s = gs - (rc ? rc[0] : 0);
    const int max_length = rc ? rc[2] : p.dyn[1];
    const int L = (rc ? rc[1] : p.dyn[0]) + s;  // rows before this append
    long long tok[8];
    for (int c = 0; c < C; ++c) tok[c] = p.raw_tokens[(long long)b * C + c];
    int n = p.needs_steps[b];
    const int u = p.unfinished[b];
    if (rc && rc[3] > 0 && L >= rc[3] && n < 0) tok[0] = p.eos_token;  // length budget reached: emit EOS on channel 0
    // wind-down trigger (:140-141)
    if (!(tok[0] >= p.speech_lo && tok[0] < p.speech_hi) && n < 0) n = C - 1;
    // teacher forcing of the delayed prompt tail (:143-145)
    if (s <= C - 2)
      for (int j = s + 1; j < C; ++j) tok[j] = p.tf_tail[((long long)b * (C - 1) + s) * C + j];
    // wind-down fill (:147-153)
    if (n > 0 && n < C - 1) {
      tok[0] = p.eos_token;
      for (int i = 1; i < C; ++i)
        if (n < C - i) tok[i] = p.pad_token;
    }
    // finished rows (:155-158)
    if (p.has_eos_criteria && !u) {
      tok[0] = p.eos_token;
      for (int i = 1; i < C; ++i) tok[i] = p.pad_token;
    }
    // append (:160) + history bitmap + next-step input
    if (L < p.max_len_rows)
      for (int c = 0; c < C; ++c) p.sequences[((long long)b * p.max_len_rows + L) * C + c] = tok[c];
    for (int c = 0; c < C; ++c) {
      p.raw_tokens[(long long)b * C + c] = tok[c];
      if (tok[c] >= 0 && tok[c] < p.cfg.vocab[c]) {
        uint32_t* w = p.seen + (long long)b * p.cfg.seen_words_per_row + p.cfg.seen_offset_words[c] + (tok[c] >> 5);
        *w |= 1u << (tok[c] & 31);  // one thread per row: no race
      }
    }
    // counters and stopping (:165-169)
    if (n > 0) n -= 1;
    const int stop = (L + 1 >= max_length) || (p.has_eos_criteria && tok[0] == p.eos_token) || (n == 0);
    int un = (u && !stop) || (n > 0);
    if (u && !un && p.finish_len[b] == 0) p.finish_len[b] = L + 1;
    p.needs_steps[b] = n;
    p.unfinished[b] = un;
    // per-row mode: a finished row's slot waits for its next request; its KV length stays put (it keeps being fed
    // [EOS, pad x7], which nothing consumes) so that it never outgrows the pages it owns
    if (!rc || u) p.positions[b] += 1;
    if (un) atomicAdd(&s_count, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.unfinished_hist[p.row_ctl ? gs % p.hist_len : gs] = s_count;
    *p.step_ptr = gs + 1;
  }
}

__global__ void init_seen_kernel(const long long* __restrict__ ids, int B, int rows, long long row_stride_b, int C,
                                 uint32_t* __restrict__ seen, const mtts_sampler_config cfg) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * rows * C;
  if (gid >= total) return;
  const int c = gid % C;
  const long long r = (gid / C) % rows;
  const long long b = gid / ((long long)C * rows);
  const long long tok = ids[b * row_stride_b + r * C + c];
  if (tok < 0 || tok >= cfg.vocab[c]) return;
  atomicOr(seen + b * cfg.seen_words_per_row + cfg.seen_offset_words[c] + (tok >> 5), 1u << (tok & 31));
}

// ------------------------------------------------------------------------------------------------
// Fused LM heads + greedy pick: the heads GEMM (gemm_tc.cu, argmax epilogue) left, per batch row and 32-row quarter of
// the stacked head matrix, the best and second-best bf16 logit as keys (order-preserving value image << 16 | 31 - lane).
// One CTA per (row, channel) scans the channel's quarters, skips the one index the reference masks at this step
// (pad for a live speech channel, EOS while the prompt tail is forced: modeling_asteroid.py:124-128) — which is why two
// candidates per quarter are enough — and keeps the best value, lowest index on ties (torch.argmax).
// ------------------------------------------------------------------------------------------------
struct PickParams {
  const uint32_t* keys;  // [B][n_quarters][2]
  int n_quarters;
  mtts_sampler_config cfg;
  const int* step_ptr;
  const int* row_ctl;
  long long* out_tokens;
};

__global__ void __launch_bounds__(256) heads_pick_kernel(const PickParams p) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, c = blockIdx.y;
  const mtts_sampler_config& cfg = p.cfg;
  const int step = *p.step_ptr - (p.row_ctl ? p.row_ctl[b * 4] : 0);
  int mask_idx = -1;
  if (c != 0 && step >= c) mask_idx = cfg.pad_token;
  if (c == 0 && step <= cfg.channels - 2) mask_idx = cfg.eos_mask_token;
  const int lo = cfg.logit_offset[c], V = cfg.vocab[c];
  const int q0 = lo >> 5, q1 = (lo + V + 31) >> 5;
  const uint32_t* kp = p.keys + ((long long)b * p.n_quarters) * 2;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = q0 * 2 + threadIdx.x; i < q1 * 2; i += 256) {
    const uint32_t key = __ldcg(kp + i);
    if (key == 0u) continue;
    const int idx = (i >> 1) * 32 + (31 - (int)(key & 31u)) - lo;
    if (idx < 0 || idx >= V || idx == mask_idx) continue;
    const uint32_t ord = key >> 16;
    const uint32_t hb = (ord & 0x8000u) ? (ord & 0x7fffu) : (~ord & 0xffffu);
    const float v = __uint_as_float(hb << 16);
    if (better(v, idx, bv, bi)) { bv = v; bi = idx; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = bv; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = threadIdx.x < 8 ? s_val[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < 8 ? s_idx[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) p.out_tokens[(long long)b * cfg.channels + c] = bi == 0x7fffffff ? 0 : bi;
  }
}

// A sampled channel wider than the candidate list is drawn by sample_exact_kernel when its top_k is beyond what the
// list's threshold search reaches, or when it has no filter at all; a nucleus without top-k goes through the list first
// and falls back per row (ws.redo).
static bool exact_only(const mtts_sampler_config* cfg, int c) {
  if (!cfg->do_sample[c] || cfg->vocab[c] <= kCap) return false;
  const int k = cfg->top_k[c] > 0 ? (cfg->top_k[c] < cfg->vocab[c] ? cfg->top_k[c] : cfg->vocab[c]) : cfg->vocab[c];
  const bool nucleus_only = cfg->top_k[c] <= 0 && cfg->has_top_p[c] && cfg->top_p[c] < 1.0f;
  return k > kThreads && !nucleus_only;
}

int validate_cfg(const mtts_sampler_config* cfg) {
  MTTS_REQUIRE(cfg != nullptr, "sampler: null config");
  MTTS_REQUIRE(cfg->channels >= 1 && cfg->channels <= 8, "sampler: channels must be in [1,8]");
  for (int c = 0; c < cfg->channels; ++c) {
    MTTS_REQUIRE(cfg->vocab[c] > 0, "sampler: vocab[%d] must be positive", c);
    if (cfg->do_sample[c]) {
      if (cfg->has_temp[c]) MTTS_REQUIRE(cfg->temperature[c] > 0.f, "sampler: temperature must be > 0");
      if (cfg->has_top_p[c]) MTTS_REQUIRE(cfg->top_p[c] >= 0.f && cfg->top_p[c] <= 1.f, "sampler: top_p must be in [0,1]");
    }
  }
  return MTTS_OK;
}

}  // namespace

extern "C" int mtts_sampler_init_history(const long long* ids, int B, int rows, long long row_stride_b,
                                         const mtts_sampler_config* cfg, uint32_t* seen, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  if (B <= 0 || rows <= 0) return MTTS_OK;
  MTTS_REQUIRE(ids && seen, "mtts_sampler_init_history: null pointer");
  const long long total = (long long)B * rows * cfg->channels;
  init_seen_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(ids, B, rows, row_stride_b, cfg->channels, seen,
                                                                         *cfg);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

// the slice kernels keep 32 KB of static shared memory; ask for the largest carve-out so that several CTAs share an SM
int mtts_configure_sampler() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(sample_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(sample_finish_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return MTTS_OK;
}

static size_t sample_ws_layout(int B, int C, SampleWs* ws, uint8_t* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const size_t bc = (size_t)B * C;
  // integer state first: [tickets | cand_count | redo] must be ZERO before the first launch (kernels leave them zero)
  int* tickets = reinterpret_cast<int*>(take(bc * 2 * sizeof(int)));
  int* cand_count = reinterpret_cast<int*>(take(bc * sizeof(int)));
  int* redo = reinterpret_cast<int*>(take(bc * sizeof(int)));
  float* slice_val = reinterpret_cast<float*>(take(bc * kMaxSlices * sizeof(float)));
  int* slice_idx = reinterpret_cast<int*>(take(bc * kMaxSlices * sizeof(int)));
  float* reported = reinterpret_cast<float*>(take((size_t)B * kMaxSlices * kReport * sizeof(float)));
  float* thr = reinterpret_cast<float*>(take(bc * 4 * sizeof(float)));
  float* cand_val = reinterpret_cast<float*>(take(bc * kCap * sizeof(float)));
  int* cand_idx = reinterpret_cast<int*>(take(bc * kCap * sizeof(int)));
  if (ws) {
    ws->tickets = tickets; ws->cand_count = cand_count; ws->redo = redo; ws->slice_val = slice_val; ws->slice_idx = slice_idx;
    ws->reported = reported; ws->thr = thr; ws->cand_val = cand_val; ws->cand_idx = cand_idx;
  }
  return off;
}

extern "C" size_t mtts_sample8_workspace_bytes(int B, int channels) {
  return sample_ws_layout(B > 0 ? B : 1, channels > 0 ? channels : 8, nullptr, nullptr);
}

extern "C" int mtts_sample8(const void* logits, long long ld, int B, const mtts_sampler_config* cfg,
                            const uint32_t* seen, const int* step_ptr, const unsigned long long* seed_ptr,
                            long long* out_tokens, int* err_flag, void* workspace, size_t workspace_bytes, void* stream_) {
  return mtts_sample8_rows(logits, ld, B, cfg, seen, step_ptr, nullptr, seed_ptr, out_tokens, err_flag, workspace,
                           workspace_bytes, stream_);
}

extern "C" int mtts_sample8_rows(const void* logits, long long ld, int B, const mtts_sampler_config* cfg,
                                 const uint32_t* seen, const int* step_ptr, const int* row_ctl,
                                 const unsigned long long* seed_ptr, long long* out_tokens, int* err_flag, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  if (B <= 0) return MTTS_OK;
  MTTS_REQUIRE(logits && seen && step_ptr && seed_ptr && out_tokens, "mtts_sample8: null pointer");
  MTTS_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, "mtts_sample8: logits rows must be 16-byte aligned");
  MTTS_REQUIRE(workspace && workspace_bytes >= mtts_sample8_workspace_bytes(B, cfg->channels),
               "mtts_sample8: workspace too small (need %zu bytes, first 4 KiB-aligned integer area zeroed once)",
               mtts_sample8_workspace_bytes(B, cfg->channels));
  SampleParams2 p;
  memset(&p, 0, sizeof(p));
  p.logits = reinterpret_cast<const bf16*>(logits); p.ld = ld; p.cfg = *cfg; p.seen = seen; p.step_ptr = step_ptr;
  p.row_ctl = row_ctl;
  p.seed_ptr = seed_ptr; p.out_tokens = out_tokens; p.err_flag = err_flag;
  sample_ws_layout(B, cfg->channels, &p.ws, reinterpret_cast<uint8_t*>(workspace));
  int total = 0;
  bool any_sample = false;
  for (int c = 0; c < cfg->channels; ++c) any_sample |= cfg->do_sample[c] != 0;
  // sampled channels keep one piece per CTA (their threshold search reports per-thread maxima of one piece)
  // batch 256: 127 us with one piece per CTA, 87 / 68 / 62 / 63 us with 2 / 4 / 8 / 16; batch 64: 34 / 24 / 20 / 19 / 20 us
  // pieces of 4096 logits per CTA: a sampled step at batch 256 is 11 520 slice CTAs whose fixed cost (launch, ticket,
  // fence) dominates; four pieces per CTA leave 10 CTAs per 152,697-way row
  int chunk = any_sample ? (B >= 64 ? 4 : B >= 16 ? 2 : 1) : (B >= 128 ? 8 : B >= 32 ? 4 : B >= 16 ? 2 : 1);
  if (const char* e = getenv("MTTS_SAMPLE_CHUNK")) {  // experiment knob (scripts/bench_sampler.py)
    if (atoi(e) >= 1 && atoi(e) <= (any_sample ? 8 : 38)) chunk = atoi(e);
  }
  p.chunk = chunk;
  int n_exact = 0;  // sampled channels wider than the candidate list: sample_exact_kernel draws them, always or per row
  for (int c = 0; c < cfg->channels; ++c) {
    if (exact_only(cfg, c)) p.exact_mask |= 1u << c;
    if (cfg->do_sample[c] && cfg->vocab[c] > kCap && (exact_only(cfg, c) || (cfg->top_k[c] <= 0 && cfg->has_top_p[c])))
      p.exact_channels[n_exact++] = (unsigned char)c;
  }
  for (int c = 0; c < cfg->channels; ++c) {
    MTTS_REQUIRE(cfg->logit_offset[c] % 8 == 0, "mtts_sample8: logit_offset[%d] must be a multiple of 8", c);
    const int S = (cfg->vocab[c] + kSlice * chunk - 1) / (kSlice * chunk);
    MTTS_REQUIRE(S <= kMaxSlices && total + S <= kMaxSlices, "mtts_sample8: vocabulary too large for %d slices", kMaxSlices);
    p.slices_of[c] = (unsigned char)S;
    for (int s2 = 0; s2 < S; ++s2) {
      p.slice_channel[total] = (unsigned char)c;
      p.slice_index[total] = (unsigned char)s2;
      ++total;
    }
  }
  p.total_slices = total;
  MTTS_CUDA_CHECK(mtts_launch(sample_scan_kernel, dim3(B, total), dim3(kThreads), 0, stream, p));
  MTTS_LAUNCH_CHECK();
  if (any_sample) {
    MTTS_CUDA_CHECK(mtts_launch(sample_finish_kernel, dim3(B, total), dim3(kThreads), 0, stream, p));
    MTTS_LAUNCH_CHECK();
  }
  if (n_exact > 0) {
    MTTS_CUDA_CHECK(mtts_launch(sample_exact_kernel, dim3(B, n_exact), dim3(kThreads), 0, stream, p));
    MTTS_LAUNCH_CHECK();
  }
  return MTTS_OK;
}

extern "C" int mtts_delay_step(long long* tokens, const long long* tf_tail, long long* sequences, long long max_len_rows,
                               int* unfinished, int* needs_steps, int* positions, uint32_t* seen, int* step_ptr,
                               int* unfinished_hist, int* finish_len, int B, const int* dyn_params,
                               int speech_lo, int speech_hi, int eos_token, int has_eos_criteria,
                               const mtts_sampler_config* cfg, void* stream_) {
  return mtts_delay_step_rows(tokens, tf_tail, sequences, max_len_rows, unfinished, needs_steps, positions, seen, step_ptr,
                              unfinished_hist, 0, finish_len, B, dyn_params, nullptr, speech_lo, speech_hi, eos_token,
                              has_eos_criteria, cfg, stream_);
}

extern "C" int mtts_delay_step_rows(long long* tokens, const long long* tf_tail, long long* sequences,
                                    long long max_len_rows, int* unfinished, int* needs_steps, int* positions,
                                    uint32_t* seen, int* step_ptr, int* unfinished_hist, int hist_len, int* finish_len,
                                    int B, const int* dyn_params, const int* row_ctl, int speech_lo, int speech_hi,
                                    int eos_token, int has_eos_criteria, const mtts_sampler_config* cfg, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  MTTS_REQUIRE(B >= 1 && B <= 1024, "mtts_delay_step: batch must be in [1,1024] (got %d)", B);
  MTTS_REQUIRE(tokens && tf_tail && sequences && unfinished && needs_steps && positions && seen && step_ptr &&
                   unfinished_hist && finish_len && (dyn_params || row_ctl),
               "mtts_delay_step: null pointer");
  MTTS_REQUIRE(row_ctl == nullptr || hist_len > 0, "mtts_delay_step_rows: hist_len must be positive with row_ctl");
  StepParams p;
  p.raw_tokens = tokens; p.tf_tail = tf_tail; p.sequences = sequences; p.max_len_rows = max_len_rows;
  p.unfinished = unfinished; p.needs_steps = needs_steps; p.positions = positions; p.seen = seen; p.step_ptr = step_ptr;
  p.unfinished_hist = unfinished_hist; p.hist_len = hist_len; p.finish_len = finish_len; p.B = B; p.C = cfg->channels;
  p.dyn = dyn_params; p.row_ctl = row_ctl;
  p.speech_lo = speech_lo; p.speech_hi = speech_hi; p.eos_token = eos_token;
  p.pad_token = cfg->pad_token; p.has_eos_criteria = has_eos_criteria; p.cfg = *cfg;
  const int threads = ((B + 31) / 32) * 32;
  MTTS_CUDA_CHECK(mtts_launch(delay_step_kernel, dim3(1), dim3(threads), 0, stream, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

// ---- fused LM heads + sampler (mtts_heads8_sample) -----------------------------------------------------
static bool heads_fusable(const mtts_sampler_config* cfg, int B) {
  if (B <= 64) return false;  // the argmax epilogue lives in the CTA-pair GEMM that batches above 64 use
  for (int c = 0; c < cfg->channels; ++c) {
    if (cfg->do_sample[c] || cfg->has_rep[c]) return false;  // a repetition penalty rescales arbitrary logits
    if (cfg->logit_offset[c] % 32 != 0) return false;
  }
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("MTTS_FUSE_HEADS");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  return off == 0;
}

extern "C" size_t mtts_heads8_sample_workspace_bytes(int B, int vpad, int channels) {
  return mtts_sample8_workspace_bytes(B, channels) + (size_t)B * (size_t)(vpad / 32 + 1) * 2 * sizeof(uint32_t) + 256;
}

extern "C" int mtts_heads8_sample_fused(const mtts_sampler_config* cfg, int B) {
  return (validate_cfg(cfg) == MTTS_OK && heads_fusable(cfg, B)) ? 1 : 0;
}

extern "C" int mtts_heads8_sample(const void* hidden, long long ld_hidden, const void* heads, long long ld_heads, int B,
                                  int hidden_size, int vpad, const mtts_sampler_config* cfg, const uint32_t* seen,
                                  const int* step_ptr, const int* row_ctl, const unsigned long long* seed_ptr,
                                  void* logits, long long ld_logits, long long* out_tokens, int* err_flag, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  if (B <= 0) return MTTS_OK;
  MTTS_REQUIRE(hidden && heads && step_ptr && out_tokens && workspace, "mtts_heads8_sample: null pointer");
  MTTS_REQUIRE(workspace_bytes >= mtts_heads8_sample_workspace_bytes(B, vpad, cfg->channels),
               "mtts_heads8_sample: workspace too small");
  const size_t sample_ws = mtts_sample8_workspace_bytes(B, cfg->channels);
  if (heads_fusable(cfg, B) && vpad % 32 == 0) {
    // greedy rows: the [B, vpad] logits are never written; the GEMM epilogue leaves 2 candidates per 32-row quarter
    uint32_t* keys = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(workspace) + sample_ws + 255) & ~uintptr_t(255));
    int lo[8], hi[8];
    for (int c = 0; c < cfg->channels; ++c) { lo[c] = cfg->logit_offset[c]; hi[c] = cfg->logit_offset[c] + cfg->vocab[c]; }
    rc = mtts_gemm_heads_argmax(hidden, ld_hidden, heads, ld_heads, B, vpad, hidden_size, cfg->channels, lo, hi, keys, stream);
    if (rc) return rc;
    PickParams pp;
    pp.keys = keys; pp.n_quarters = vpad / 32; pp.cfg = *cfg; pp.step_ptr = step_ptr; pp.row_ctl = row_ctl;
    pp.out_tokens = out_tokens;
    MTTS_CUDA_CHECK(mtts_launch(heads_pick_kernel, dim3(B, cfg->channels), dim3(256), 0, stream, pp));
    MTTS_LAUNCH_CHECK();
    return MTTS_OK;
  }
  MTTS_REQUIRE(logits != nullptr && seen != nullptr && seed_ptr != nullptr, "mtts_heads8_sample: the unfused path needs logits / seen / seed");
  rc = mtts_gemm(hidden, ld_hidden, heads, ld_heads, logits, ld_logits, B, vpad, hidden_size, MTTS_DTYPE_BF16, MTTS_DTYPE_BF16, 0,
                 nullptr, nullptr, nullptr, 0, nullptr, 0, stream_);
  if (rc) return rc;
  return mtts_sample8_rows(logits, ld_logits, B, cfg, seen, step_ptr, row_ctl, seed_ptr, out_tokens, err_flag, workspace,
                           sample_ws, stream_);
}
