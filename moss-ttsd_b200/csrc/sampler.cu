// Delay-pattern sampler: per-channel logit masks + HF logits processors + draw, and the per-row state
// machine of CustomMixin._sample, all on the device (the reference needs ~60 launches and two host syncs
// per step for this, modeling_asteroid.py:123-169).
//
//   sample8_kernel     one CTA per (row, channel): masks (modeling_asteroid.py:124-128), repetition penalty ->
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

namespace {

constexpr int kThreads = 512;
constexpr int kCap = 2048;  // candidate list capacity

struct SampleParams {
  const bf16* logits;
  long long ld;
  mtts_sampler_config cfg;
  const uint32_t* seen;
  const int* step_ptr;
  unsigned long long seed;
  long long* out_tokens;  // [B, channels]
  int* err_flag;
};

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

__global__ void __launch_bounds__(kThreads) sample8_kernel(const SampleParams p) {
  __shared__ float s_val[kCap];
  __shared__ int s_idx[kCap];
  __shared__ float s_red[33];
  __shared__ int s_count;
  __shared__ float s_thr;
  __shared__ int s_keep;
  __shared__ float s_zkeep;

  const int b = blockIdx.x, c = blockIdx.y;
  const mtts_sampler_config& cfg = p.cfg;
  const int V = cfg.vocab[c];
  const int step = *p.step_ptr;
  ScoreCtx sc;
  sc.lg = p.logits + (long long)b * p.ld + cfg.logit_offset[c];
  sc.seen = p.seen + (long long)b * cfg.seen_words_per_row + cfg.seen_offset_words[c];
  sc.mask_idx = -1;
  if (c != 0 && step >= c) sc.mask_idx = cfg.pad_token;             // channel c is live: pad is illegal
  if (c == 0 && step <= cfg.channels - 2) sc.mask_idx = cfg.eos_mask_token;  // no EOS while the prompt tail is forced
  sc.has_rep = cfg.has_rep[c] != 0;
  sc.pen = cfg.rep_penalty[c];
  sc.has_temp = cfg.has_temp[c] != 0;
  sc.temp = cfg.temperature[c];
  const int tid = threadIdx.x;

  // ---------------- pass 1: per-thread maximum (also the greedy answer)
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = tid; j < V; j += kThreads) {
    const float s = score_at(sc, j);
    if (better(s, j, bv, bi)) { bv = s; bi = j; }
  }
  if (!cfg.do_sample[c]) {
    // block argmax, lowest index on ties (torch.argmax)
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
      if (tid == 0) p.out_tokens[(long long)b * cfg.channels + c] = bi == 0x7fffffff ? 0 : bi;
    }
    return;
  }

  // ---------------- candidate threshold
  int k = cfg.top_k[c] > 0 ? min(cfg.top_k[c], V) : V;
  if (k <= kThreads && V > kCap) {
    s_val[tid] = bv;
    s_idx[tid] = bi;
    __syncthreads();
    bitonic_sort_desc(s_val, s_idx, kThreads);
    if (tid == 0) s_thr = s_val[k - 1];  // >= k scores are >= this value
    __syncthreads();
  } else {
    if (tid == 0) s_thr = -INFINITY;
    __syncthreads();
  }
  const float thr = s_thr;
  if (tid == 0) s_count = 0;
  __syncthreads();

  // ---------------- pass 2: collect candidates
  for (int j = tid; j < V; j += kThreads) {
    const float s = score_at(sc, j);
    if (s >= thr) {
      const int pos = atomicAdd(&s_count, 1);
      if (pos < kCap) { s_val[pos] = s; s_idx[pos] = j; }
    }
  }
  __syncthreads();
  int n = s_count;
  if (n > kCap) {
    if (tid == 0 && p.err_flag) *p.err_flag = 3;  // candidate overflow (pathological ties); truncated
    n = kCap;
  }
  int npad = 1;
  while (npad < n) npad <<= 1;
  for (int t = n + tid; t < npad; t += kThreads) { s_val[t] = -INFINITY; s_idx[t] = -1; }
  __syncthreads();
  bitonic_sort_desc(s_val, s_idx, npad);

  // ---------------- top-k (ties with the k-th value are kept, HF: scores < kth -> -inf), top-p, draw
  if (tid < 32) {
    const int lane = tid;
    int nk = n;
    if (cfg.top_k[c] > 0 && k < n) {
      const float kth = s_val[k - 1];
      // first position whose value is < kth (list is sorted descending)
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
    // probabilities (unnormalised), stored in place
    float part = 0.f;
    const int chunk = (nk + 31) / 32;
    const int t0 = lane * chunk, t1 = min(nk, t0 + chunk);
    for (int t = t0; t < t1; ++t) {
      const float e = expf(s_val[t] - vmax);
      s_val[t] = e;
      part += e;
    }
    // inclusive scan of per-lane partial sums (descending order)
    float incl = part;
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const float Z = __shfl_sync(0xffffffffu, incl, 31);
    int keep = nk;
    if (cfg.has_top_p[c]) {
      // HF: sort ascending, cum = cumsum(softmax); remove cum <= 1 - top_p; always keep the largest.
      // In descending order: position t is removed iff (mass of t and everything after it) / Z <= 1 - top_p.
      const float limit = (1.0f - cfg.top_p[c]);
      float before = incl - part;  // mass strictly before this lane's chunk
      int my_keep = 0;
      for (int t = t0; t < t1; ++t) {
        const float tail = (Z - before) / Z;  // mass of t..end
        if (t == 0 || tail > limit) my_keep = t + 1;
        before += s_val[t];
      }
      for (int o = 16; o > 0; o >>= 1) my_keep = max(my_keep, __shfl_xor_sync(0xffffffffu, my_keep, o));
      keep = max(my_keep, 1);
    }
    // mass of the kept prefix
    float kpart = 0.f;
    for (int t = t0; t < min(t1, keep); ++t) kpart += s_val[t];
    float kincl = kpart;
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, kincl, o);
      if (lane >= o) kincl += up;
    }
    const float Zk = __shfl_sync(0xffffffffu, kincl, 31);
    const float u = philox_uniform(p.seed, (uint32_t)step, (uint32_t)(b * 8 + c));
    const float target = u * Zk;
    // the lane whose chunk contains the target walks it
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
    if (lane == 0) p.out_tokens[(long long)b * cfg.channels + c] = s_idx[choice];
  }
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
  int* finish_len;           // [B] sequence length (rows) at which the row finished (0 = not yet)
  int B, C, P, max_length;
  int speech_lo, speech_hi, eos_token, pad_token, has_eos_criteria;
  mtts_sampler_config cfg;
};

__global__ void delay_step_kernel(const StepParams p) {
  __shared__ int s_count;
  const int b = threadIdx.x;
  const int s = *p.step_ptr;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (b < p.B) {
    const int C = p.C;
    const int L = p.P + s;  // rows before this append
    long long tok[8];
    for (int c = 0; c < C; ++c) tok[c] = p.raw_tokens[(long long)b * C + c];
    int n = p.needs_steps[b];
    const int u = p.unfinished[b];
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
    const int stop = (L + 1 >= p.max_length) || (p.has_eos_criteria && tok[0] == p.eos_token) || (n == 0);
    int un = (u && !stop) || (n > 0);
    if (u && !un && p.finish_len[b] == 0) p.finish_len[b] = L + 1;
    p.needs_steps[b] = n;
    p.unfinished[b] = un;
    p.positions[b] += 1;
    if (un) atomicAdd(&s_count, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.unfinished_hist[s] = s_count;
    *p.step_ptr = s + 1;
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

int validate_cfg(const mtts_sampler_config* cfg) {
  MTTS_REQUIRE(cfg != nullptr, "sampler: null config");
  MTTS_REQUIRE(cfg->channels >= 1 && cfg->channels <= 8, "sampler: channels must be in [1,8]");
  for (int c = 0; c < cfg->channels; ++c) {
    MTTS_REQUIRE(cfg->vocab[c] > 0, "sampler: vocab[%d] must be positive", c);
    if (cfg->do_sample[c]) {
      const int k = cfg->top_k[c] > 0 ? (cfg->top_k[c] < cfg->vocab[c] ? cfg->top_k[c] : cfg->vocab[c]) : cfg->vocab[c];
      if (cfg->vocab[c] > kCap && k > kThreads)
        return mtts_set_error(MTTS_ERR_UNSUPPORTED,
                              "sampler: channel %d samples over %d tokens with top_k=%d; this build needs top_k <= %d "
                              "(or vocab <= %d) for sampled channels",
                              c, cfg->vocab[c], cfg->top_k[c], kThreads, kCap);
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

extern "C" int mtts_sample8(const void* logits, long long ld, int B, const mtts_sampler_config* cfg,
                            const uint32_t* seen, const int* step_ptr, unsigned long long seed, long long* out_tokens,
                            int* err_flag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  if (B <= 0) return MTTS_OK;
  MTTS_REQUIRE(logits && seen && step_ptr && out_tokens, "mtts_sample8: null pointer");
  SampleParams p;
  p.logits = reinterpret_cast<const bf16*>(logits); p.ld = ld; p.cfg = *cfg; p.seen = seen; p.step_ptr = step_ptr;
  p.seed = seed; p.out_tokens = out_tokens; p.err_flag = err_flag;
  sample8_kernel<<<dim3(B, cfg->channels), kThreads, 0, stream>>>(p);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

extern "C" int mtts_delay_step(long long* tokens, const long long* tf_tail, long long* sequences, long long max_len_rows,
                               int* unfinished, int* needs_steps, int* positions, uint32_t* seen, int* step_ptr,
                               int* unfinished_hist, int* finish_len, int B, int prompt_rows, int max_length,
                               int speech_lo, int speech_hi, int eos_token, int has_eos_criteria,
                               const mtts_sampler_config* cfg, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  MTTS_REQUIRE(B >= 1 && B <= 1024, "mtts_delay_step: batch must be in [1,1024] (got %d)", B);
  MTTS_REQUIRE(tokens && tf_tail && sequences && unfinished && needs_steps && positions && seen && step_ptr &&
                   unfinished_hist && finish_len,
               "mtts_delay_step: null pointer");
  StepParams p;
  p.raw_tokens = tokens; p.tf_tail = tf_tail; p.sequences = sequences; p.max_len_rows = max_len_rows;
  p.unfinished = unfinished; p.needs_steps = needs_steps; p.positions = positions; p.seen = seen; p.step_ptr = step_ptr;
  p.unfinished_hist = unfinished_hist; p.finish_len = finish_len; p.B = B; p.C = cfg->channels; p.P = prompt_rows;
  p.max_length = max_length; p.speech_lo = speech_lo; p.speech_hi = speech_hi; p.eos_token = eos_token;
  p.pad_token = cfg->pad_token; p.has_eos_criteria = has_eos_criteria; p.cfg = *cfg;
  const int threads = ((B + 31) / 32) * 32;
  delay_step_kernel<<<1, threads, 0, stream>>>(p);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
