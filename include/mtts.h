/*
 * libmtts — C-ABI of the B200-native MOSS-TTSD generation hot path.
 *
 * The reference (zsc/MOSS-TTSD) has no FFI layer: its hot path is three Python call sites
 *   model.generate(...)            generation_utils.py:406-409  -> modeling_asteroid.py:52-197 (+ HF Qwen3Model)
 *   spt.encode([wav])              generation_utils.py:198      -> XY_Tokenizer/xy_tokenizer/model.py:130-192
 *   spt.decode(codes, overlap=10)  generation_utils.py:449      -> XY_Tokenizer/xy_tokenizer/model.py:194-256
 * Each entry point below replaces the stock torch ops behind one of those call sites; the reference
 * file:line it stands in for is cited on every declaration. INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device;
 *   - nothing allocates or frees caller-visible memory: scratch comes in through `workspace` arguments whose
 *     sizes the *_workspace_bytes() queries return;
 *   - return 0 on success, <0 on error, message via mtts_last_error() (thread-local);
 *   - launches go to the CURRENT device; the library keeps no mutable global state except caches of
 *     immutable TMA descriptors.
 */
#ifndef MTTS_H_
#define MTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTTS_VERSION 100

/* element types */
#define MTTS_DTYPE_BF16 0
#define MTTS_DTYPE_F32 1
#define MTTS_DTYPE_F16 2 /* mtts_gemm operands / GELU output of the codec fp16-operand path */

/* GEMM epilogue flags (bitmask). Order of application:
 *   v = acc; +bias[n]; gelu; round to out dtype; *gamma[n]; residual[m,n] + v.
 * SWIGLU is exclusive: weight rows interleaved (2j = gate_j, 2j+1 = up_j), out[m, j] = silu(g) * u. */
#define MTTS_EPI_BIAS 1
#define MTTS_EPI_GELU 2
#define MTTS_EPI_GAMMA 4
#define MTTS_EPI_RESIDUAL 8
#define MTTS_EPI_SWIGLU 16
#define MTTS_EPI_EXACT_ACT 32 /* fp32 outputs: erff() GELU instead of the 1.5e-7 polynomial (exact-mode codec encode) */

const char* mtts_last_error(void);
int mtts_version(void);
/* Caches per-device attributes for the current device. Safe to call repeatedly. */
int mtts_init(void);
/* Number of kernels this library has launched (or recorded into a CUDA graph under capture) in this process. */
long long mtts_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense projections (tcgen05 / TMEM / TMA).   out[M,N] = epi(x[M,K] . w[N,K]^T)
 * Replaces aten::linear behind HF Qwen3Attention / Qwen3MLP (invoked from modeling_asteroid.py:273-284),
 * the 8 lm_heads (modeling_asteroid.py:412) and the codec nn.Linear / ConvTranspose1d layers
 * (XY_Tokenizer/xy_tokenizer/nn/modules.py:84-87,181-182,494-500,1111-1115,957).
 * in_dtype BF16: bf16 operands; F32: fp32 storage multiplied as TF32. fp32 accumulation either way.
 * bias/gamma are fp32 [N]; residual has the output dtype. Row strides in ELEMENTS.
 * workspace: unused since split-K partials moved to distributed shared memory (kept for ABI stability; may be NULL).
 * ---------------------------------------------------------------------------------------------- */
size_t mtts_gemm_workspace_bytes(int M, int N, int K, int in_dtype);
int mtts_gemm(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo, int M, int N,
              int K, int in_dtype, int out_dtype, int flags, const float* bias, const float* gamma,
              const void* residual, long long ldr, void* workspace, size_t workspace_bytes, void* stream);

/* Split-K projection for the decode step (1 <= M <= 256, bf16) with the reduction moved into the consumer: every CTA
 * owns all M rows of one (128-row weight tile, k-slice) and stores its fp32 partial tile into `partials`
 * [splits][M][N] (no cluster, no barrier); mtts_splitk_reduce / mtts_splitk_reduce_rmsnorm sum the slices in ascending
 * order. Replaces `aten::linear` of q/k/v, o_proj and down_proj in HF Qwen3Attention / Qwen3MLP as invoked from
 * modeling_asteroid.py:226,273-284 together with the residual add + RMSNorm that follow them
 * (modeling_qwen3.py:50-66,327-333). */
int mtts_gemm_splitk_splits(int M, int N, int K);
size_t mtts_gemm_splitk_workspace_bytes(int M, int N, int K);
int mtts_gemm_splitk(const void* x, long long ldx, const void* w, long long ldw, float* partials, size_t partial_bytes,
                     int M, int N, int K, int* splits_out, void* stream);
/* out[M, N] bf16 = bf16(sum_s partials[s]) */
int mtts_splitk_reduce(const float* partials, int splits, int M, int N, void* out, long long ldo, void* stream);
/* x[M, N] bf16 (residual stream, in place) <- bf16(x + bf16(sum_s partials[s]));  xn <- RMSNorm(x, eps) * norm_w */
int mtts_splitk_reduce_rmsnorm(const float* partials, int splits, int M, int N, void* x, long long ldx, const void* norm_w,
                               void* xn, long long ldxn, float eps, void* stream);

/* Exact-fp32 CUDA-core GEMM with generic strides (used where TF32 would break bit-exact parity, i.e. the
 * RVQ input projection quantizer.py:224,245, and as the in-library cross-check of mtts_gemm).
 *   x element (m, k) at x[(m / rows_per_batch) * x_batch_stride + (m % rows_per_batch) * x_row_stride + k * x_k_stride]
 * Same epilogue flags (no SWIGLU). in_dtype selects bf16 or fp32 operands (both accumulate in fp32 FMA). */
int mtts_gemm_simt(const void* x, int rows_per_batch, long long x_batch_stride, long long x_row_stride,
                   long long x_k_stride, const void* w, long long ldw, void* out, long long ldo, int M, int N,
                   int K, int in_dtype, int out_dtype, int flags, const float* bias, const float* gamma,
                   const void* residual, long long ldr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ResidualVQ (XY_Tokenizer/xy_tokenizer/nn/quantizer.py)
 * ---------------------------------------------------------------------------------------------- */
/* ||c||^2 per code, fp32, as `self.codebook.float().pow(2).sum(1)` (quantizer.py:169). norms: [nq, K]. */
int mtts_rvq_codebook_norms(const float* codebooks, int nq, int codebook_size, int dim, float* norms, void* stream);

/* Nearest-code search over nq residual layers (VectorQuantize.forward quantizer.py:167-172 inside
 * ResidualVQ.forward :277-327, inference branch).
 *   z         [N, dim] fp32, token-major (one row per (b, t))
 *   valid     [N] uint8 or NULL: 0 rows quantise the zero vector and contribute nothing (the `mask` of :250,278,309)
 *   codebooks [nq, K, dim] fp32, norms [nq, K]
 *   codes     [nq, N] int64 out (all_indices)
 *   zq        [N, dim] fp32 out or NULL: sum of selected code vectors over valid rows (quantized_out before output_proj)
 *   residual_out [N, dim] fp32 or NULL: final residual (test hook)
 * dim must be 512-or-less and a multiple of 4; K a multiple of 128. */
int mtts_rvq_encode(const float* z, const uint8_t* valid, const float* codebooks, const float* norms, int N, int nq,
                    int codebook_size, int dim, long long* codes, float* zq, float* residual_out, void* stream);

/* Codebook gather + sum (ResidualVQ.decode_codes quantizer.py:345-361, VectorQuantize.decode_code :193-194).
 *   codes [nq, N] int64 (row stride codes_ld), out [N, dim] fp32 token-major = sum_i codebooks[i][codes[i][n]].
 * Out-of-range codes set *err_flag (device int, may be NULL) to 1 and contribute zero. */
int mtts_rvq_decode(const long long* codes, long long codes_ld, const float* codebooks, int N, int nq,
                    int codebook_size, int dim, float* out, int* err_flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Qwen3-style decoder step (bf16).  Reference: HF Qwen3Model invoked at modeling_asteroid.py:226,273-284,
 * arithmetic and rounding points per SURVEY.md Appendix B.
 * ---------------------------------------------------------------------------------------------- */
/* out[r,:] = sum_c bf16 tables[c][ids[r,c],:] with a bf16 rounding after every add
 * (AsteroidTTSModel._prepare_multi_modal_inputs, modeling_asteroid.py:235-250).
 * ids [rows, channels] int64; tables_host: HOST array of `channels` device pointers ([vocab_c, hidden] bf16);
 * out [rows, hidden] bf16. Out-of-range ids set *err_flag = 1 and contribute zero. */
int mtts_embed_sum8(const long long* ids, int rows, int channels, const void* const* tables_host,
                    const int* vocab_sizes_host, int hidden, void* out, int* err_flag, void* stream);

/* Qwen3RMSNorm: fp32 mean-square, cast to bf16, then bf16 multiply by w. x/out [rows, hidden] bf16. */
int mtts_rmsnorm(const void* x, long long ldx, const void* w, void* out, long long ldo, int rows, int hidden, float eps,
                 void* stream);

/* Per-head RMSNorm of q and k, rotate-half RoPE, and the KV-cache append (replaces DynamicCache.update's torch.cat).
 *   qkv [rows, (Hq+2Hkv)*128] bf16; q_out [rows, Hq*128]; pools [num_pages, Hkv, page_size, 128] bf16;
 *   positions[r] = RoPE position = slot in the sequence's cache; row_seq[r] = sequence (NULL: r);
 *   block_table [num_seqs, max_pages] int32 or NULL for a contiguous cache (page = seq*max_pages + pos/page_size);
 *   inv_freq [64] fp32 = theta^(-2i/128). */
int mtts_qknorm_rope_kvappend(const void* qkv, long long ld_qkv, const void* q_norm_w, const void* k_norm_w,
                              const float* inv_freq, const int* positions, const int* row_seq, void* q_out,
                              void* k_pool, void* v_pool, const int* block_table, int max_pages, int page_size,
                              int num_pages, int rows, int num_q_heads, int num_kv_heads, int head_dim, float eps,
                              int* err_flag, void* stream);

/* Causal GQA attention over the paged/contiguous cache; decode (rows_per_tile = 1, tile arrays NULL) and prefill
 * (rows_per_tile = 64: bf16 tensor-core tiles, or 4: CUDA-core tiles; tile t covers query rows tile_row0[t] .. +tile_nrows[t]
 * of ONE sequence, consecutive positions).
 * Row r attends keys 0..positions[r] of sequence row_seq[r]. out [rows, Hq*128] bf16. nsplit > 1 splits the keys
 * across CTAs (small batches); workspace from mtts_gqa_attention_workspace_bytes, first 64 KiB zero before first use. */
size_t mtts_gqa_attention_workspace_bytes(int tiles, int num_kv_heads, int group, int rows_per_tile, int nsplit);
int mtts_gqa_attention(const void* q, const void* k_pool, const void* v_pool, const int* block_table, int max_pages,
                       int page_size, const int* tile_row0, const int* tile_nrows, const int* row_seq,
                       const int* positions, void* out, int tiles, int rows_per_tile, int num_q_heads,
                       int num_kv_heads, int head_dim, int nsplit, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Delay-pattern sampler (CustomMixin._sample, modeling_asteroid.py:52-197).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mtts_sampler_config {
  int channels;            /* 8 */
  int vocab[8];            /* logits per channel: 152697, 1025 x7 */
  int logit_offset[8];     /* column of channel c's first logit in the fused-head output row */
  int do_sample[8];        /* generation_config.do_samples */
  int has_rep[8];   float rep_penalty[8];   /* RepetitionPenaltyLogitsProcessor */
  int has_temp[8];  float temperature[8];   /* TemperatureLogitsWarper */
  int top_k[8];                             /* TopKLogitsWarper, 0 = absent */
  int has_top_p[8]; float top_p[8];         /* TopPLogitsWarper */
  int seen_offset_words[8];  /* per-channel offset (32-bit words) inside one row of the history bitmap */
  int seen_words_per_row;
  int pad_token;           /* 1024    (modeling_asteroid.py:126) */
  int eos_mask_token;      /* 152694  (modeling_asteroid.py:128) */
} mtts_sampler_config;

/* Mark every token of ids[:, :rows, c] (int64 [B, *, channels], batch stride row_stride_b elements) in the
 * per-(row, channel) history bitmap `seen` (zeroed by the caller). History includes left-pad rows, as in the reference. */
int mtts_sampler_init_history(const long long* ids, int B, int rows, long long row_stride_b,
                              const mtts_sampler_config* cfg, uint32_t* seen, void* stream);

/* One draw per (row, channel) from the bf16 fused-head logits [B, ld]: masks -> repetition penalty -> temperature ->
 * top-k -> top-p -> multinomial (Philox, stream = (*seed_ptr, step, row, channel); the seed lives in device memory so
 * that a captured CUDA graph can be replayed with a new seed) or argmax. out_tokens [B, channels] int64.
 * *step_ptr is the device-resident step counter s (0 = first generated row). `workspace`: mtts_sample8_workspace_bytes()
 * bytes, zero-filled once by the caller (the kernels leave it clean); logits rows and per-channel offsets 16-byte aligned.
 * Every HF combination is accepted: a channel wider than the 2048-entry candidate list (the 152,697-way text channel) with
 * top_k > 512, with no filter at all, or with a nucleus that outgrows the list is drawn by an exact multi-pass kernel
 * (radix-select top-k, bisection top-p, inverse CDF) instead of the candidate-list fast path. */
size_t mtts_sample8_workspace_bytes(int B, int channels);
int mtts_sample8(const void* logits, long long ld, int B, const mtts_sampler_config* cfg, const uint32_t* seen,
                 const int* step_ptr, const unsigned long long* seed_ptr, long long* out_tokens, int* err_flag,
                 void* workspace, size_t workspace_bytes, void* stream);

/* The 8 LM heads FUSED with the sampler (replaces `lm_heads[i](hidden)` modeling_asteroid.py:412 + the mask /
 * processor / draw block :123-138 for the last position). hidden [B, hidden_size] bf16 (final-norm output), heads
 * [vpad, hidden_size] bf16 (the 8 head matrices stacked, each padded to a multiple of 32 rows). When every channel is
 * greedy without a repetition penalty and B > 64, the [B, vpad] logits are never written: the heads GEMM's epilogue
 * reduces every 32-row quarter to its best and second-best bf16 logit (warp-level integer max over order-preserving
 * keys) and a (row, channel) pick kernel applies the step's pad / EOS mask and takes the argmax (lowest index on
 * ties). Otherwise it is mtts_gemm -> `logits` -> mtts_sample8_rows. mtts_heads8_sample_fused() tells which. */
size_t mtts_heads8_sample_workspace_bytes(int B, int vpad, int channels);
int mtts_heads8_sample_fused(const mtts_sampler_config* cfg, int B);
int mtts_heads8_sample(const void* hidden, long long ld_hidden, const void* heads, long long ld_heads, int B, int hidden_size,
                       int vpad, const mtts_sampler_config* cfg, const uint32_t* seen, const int* step_ptr, const int* row_ctl,
                       const unsigned long long* seed_ptr, void* logits, long long ld_logits, long long* out_tokens,
                       int* err_flag, void* workspace, size_t workspace_bytes, void* stream);

/* The per-row state machine after the draw: wind-down trigger, teacher forcing (tf_tail [B, channels-1, channels] =
 * prompt[:, P:P+channels-1, :]), wind-down fill, finished fill, append to sequences [B, max_len_rows, channels] at row
 * P + s (P = dyn_params[0], max_length = dyn_params[1]; device ints, so a captured graph survives a new prompt), history bitmap update, counters/stopping, positions[b] += 1, unfinished_hist[s] = #unfinished rows,
 * finish_len[b] = length at which row b finished, and finally *step_ptr += 1. tokens [B, channels] is updated in place
 * and is the next step's input_ids. B <= 1024. */
int mtts_delay_step(long long* tokens, const long long* tf_tail, long long* sequences, long long max_len_rows,
                    int* unfinished, int* needs_steps, int* positions, uint32_t* seen, int* step_ptr,
                    int* unfinished_hist, int* finish_len, int B, const int* dyn_params, int speech_lo, int speech_hi,
                    int eos_token, int has_eos_criteria, const mtts_sampler_config* cfg, void* stream);

/* Per-row variants (continuous batching): `row_ctl` [B][4] int32 = {step0, P, max_length, eos_at} gives every
 * row its own step origin (its state machine runs at step - step0), prompt length and max_length, so a finished row's
 * slot can be refilled with a queued request while the other rows keep decoding (the reference keeps feeding finished
 * rows [EOS, pad x7] until the longest row is done, modeling_asteroid.py:155-158,166-169). eos_at > 0: channel 0 is
 * forced to EOS from sequence row eos_at on (per-request length budget; the row winds down as after a sampled EOS).
 * `unfinished_hist` is a ring of `hist_len` entries here. row_ctl == NULL: identical to the functions above. */
int mtts_sample8_rows(const void* logits, long long ld, int B, const mtts_sampler_config* cfg, const uint32_t* seen,
                      const int* step_ptr, const int* row_ctl, const unsigned long long* seed_ptr, long long* out_tokens,
                      int* err_flag, void* workspace, size_t workspace_bytes, void* stream);
int mtts_delay_step_rows(long long* tokens, const long long* tf_tail, long long* sequences, long long max_len_rows,
                         int* unfinished, int* needs_steps, int* positions, uint32_t* seen, int* step_ptr,
                         int* unfinished_hist, int hist_len, int* finish_len, int B, const int* dyn_params,
                         const int* row_ctl, int speech_lo, int speech_hi, int eos_token, int has_eos_criteria,
                         const mtts_sampler_config* cfg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * XY_Tokenizer decode path, fp32, token-major activations [batch*frames, channels]
 * (XY_Tokenizer/xy_tokenizer/model.py:103-128 -> nn/modules.py). All dense layers go through mtts_gemm (TF32).
 * ---------------------------------------------------------------------------------------------- */
/* nn.LayerNorm over C; rows at or beyond lengths[row / rows_per_item] are written as zeros when lengths != NULL
 * (modules.py:171,182,554 and the masking of :407,626). */
int mtts_layernorm(const float* x, const float* w, const float* b, float* out, long long rows, int C, float eps,
                   const int* lengths, int rows_per_item, void* stream);

/* fp16-operand path of the decoder's large GEMMs (mtts_gemm with MTTS_DTYPE_F16 operands: the 10-bit mantissa of the TF32
 * path at twice the tensor rate and half the operand bytes): the normalisations write their output — the next GEMM's
 * activation operand — directly as fp16. Same arithmetic as mtts_layernorm / mtts_dwconv7_ln, one rounding at the store. */
int mtts_layernorm_f16(const float* x, const float* w, const float* b, void* out_f16, long long rows, int C, float eps,
                       const int* lengths, int rows_per_item, void* stream);
int mtts_dwconv7_ln_f16(const float* x, const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                        void* out_f16, int B, int T, int C, float eps, void* stream);

/* VarLenAttention core (modules.py:117-160), non-causal, head_dim 64: qkv [B*T, 3*H*64] (q|k|v incl. biases, q NOT yet
 * scaled), out [B*T, H*64]; keys >= lengths[b] are masked. */
int mtts_mha_varlen(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads, int head_dim,
                    void* stream);

/* Same contraction with fp16 tensor-core operands (Q, K, V, probabilities; fp32 accumulate and softmax): the attention of
 * the decoder's fp16-operand path. */
int mtts_mha_varlen_f16(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads, int head_dim,
                        void* stream);

/* The same contraction on the 5th-gen tensor cores (tcgen05.mma with S and O in TMEM, Q / K / V tiles by TMA, one thread
 * per query row for the softmax), fp16 in and out: qkv [B*T, 3*H*64] fp16, out [B*T, H*64] fp16. The attention of the
 * decoder's fp16-operand path (modules.py:117-160 as called from model.py:103-128). */
int mtts_mha_varlen_tc(const void* qkv_f16, void* out_f16, const int* lengths, int B, int T, int num_heads, int head_dim,
                       void* stream);

/* Same contraction with fp32 CUDA-core products, fp32 softmax and expf(): the exact mode of XY_Tokenizer.encode, whose
 * integer codes must match the reference's fp32 matmuls (model.py:54-101 -> modules.py:117-160). */
int mtts_mha_varlen_fp32(const float* qkv, float* out, const int* lengths, int B, int T, int num_heads, int head_dim,
                         void* stream);

/* Window bookkeeping of XY_Tokenizer.encode / decode (model.py:214-216,241-243): dst[b, i] = i < lens[b] ? src[b, i] : 0
 * for b < B, i < n; elements of 4 (waveform fp32) or 8 bytes (int64 codes); row strides in elements. */
int mtts_rows_prefix_copy(const void* src, long long lds, void* dst, long long ldd, const int* lens, int B, int n,
                          int elem_bytes, void* stream);

/* 3xTF32 operand split for fp32-accurate products on the tcgen05 TF32 path: x [rows, K] fp32 -> out [rows, 3K] with
 * hi = rna_tf32(x), lo = x - hi laid out [hi | lo | hi] (weights_order = 0, activations) or [hi | hi | lo]
 * (weights_order = 1), so that mtts_gemm over K' = 3K computes hi*hi + lo*hi + hi*lo in fp32 accumulators — the
 * fp32 `aten::linear` of the codec encoder (modules.py:84-87,181-182; `matmul.allow_tf32 = False`). */
int mtts_split_tf32x3(const float* x, long long ldx, float* out, long long ldo, long long rows, int K, int weights_order,
                      void* stream);

/* ConvNeXtBlock front half (modules.py:1142-1150): depthwise Conv1d(k=7,pad=3) + LayerNorm(C, eps). x,out [B,T,C];
 * conv_w [C,7]; zero padding at both ends of every item. */
int mtts_dwconv7_ln(const float* x, const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                    float* out, int B, int T, int C, float eps, void* stream);

/* Tap overlap-add of a ConvTranspose1d computed as GEMM (modules.py:354-368,413-419): y [B,Tin,K,Cout] ->
 * out[b,u,co] = act(bias[co] + sum_j y[b,(u-j)/stride,j,co]), u < Tout (trim), act = exact-erf GELU if gelu. */
int mtts_convt_gather(const float* y, const float* bias, float* out, int B, int Tin, int Cout, int K, int stride,
                      int Tout, int gelu, void* stream);

/* Prefill attention on the 5th-gen tensor cores (tcgen05.mma with S and O in TMEM, Q tiles and the K/V tiles of the paged
 * pool by TMA): the same contraction as mtts_gqa_attention for packed prompt rows (HF Qwen3Attention.forward, installed
 * modeling_qwen3.py:236-288 as invoked from modeling_asteroid.py:226,273-284), causal, GQA, head_dim 128. Tiles are up to
 * 128 consecutive rows of ONE sequence (tile_row0 / tile_nrows [tiles]); q [rows, Hq * 128] bf16; the keys of every row,
 * incl. the rows of this call, are already in the pools (mtts_qknorm_rope_kvappend runs first); page_size >= 64. */
int mtts_gqa_prefill_tc(const void* q, long long rows, const void* k_pool, const void* v_pool, const int* block_table,
                        int max_pages, int page_size, int num_pages, const int* tile_row0, const int* tile_nrows,
                        const int* row_seq, const int* positions, void* out, int tiles, int num_q_heads, int num_kv_heads,
                        int head_dim, void* stream);

/* Decode attention fused with its prologue: for ONE new row per sequence it does what mtts_qknorm_rope_kvappend +
 * mtts_gqa_attention do (HF Qwen3Attention.forward, installed modeling_qwen3.py:236-288: q/k RMSNorm, RoPE, cache
 * update, attention), bit-identically, in one launch: q and the new K/V row never travel through global memory and the
 * first K/V tile is requested before the kernel waits for the q/k/v projection (programmatic dependent launch).
 * qkv [rows, (Hq + 2 Hkv) * 128] bf16 is the projection output; out [rows, Hq * 128] bf16; the new key/value is
 * appended at positions[row]; workspace as for mtts_gqa_attention (tiles = rows, rows_per_tile = 1). */
int mtts_gqa_decode_fused(const void* qkv, long long ld_qkv, const void* q_norm_w, const void* k_norm_w,
                          const float* inv_freq, float eps, void* k_pool, void* v_pool, const int* block_table,
                          int max_pages, int page_size, int num_pages, const int* positions, void* out, int rows,
                          int num_q_heads, int num_kv_heads, int head_dim, int nsplit, void* workspace,
                          size_t workspace_bytes, int* err_flag, void* stream);
/* Same, with q/k/v taken as the fp32 split-K partial tiles [splits][rows][(Hq + 2 Hkv) * 128] of mtts_gemm_splitk: the
 * slices are summed in ascending order and rounded to bf16 once inside the prologue (== mtts_splitk_reduce + the call
 * above, one launch and one 2 MB round trip fewer per layer). */
int mtts_gqa_decode_fused_splitk(const float* qkv_partials, int qkv_splits, const void* q_norm_w, const void* k_norm_w,
                                 const float* inv_freq, float eps, void* k_pool, void* v_pool, const int* block_table,
                                 int max_pages, int page_size, int num_pages, const int* positions, void* out, int rows,
                                 int num_q_heads, int num_kv_heads, int head_dim, int nsplit, void* workspace,
                                 size_t workspace_bytes, int* err_flag, void* stream);

/* ---- Small-batch decode step as one persistent kernel (batch 1..4; hidden 2048, intermediate 6144, 16/8 heads x 128).
 * Runs, for ONE new row per sequence, everything between the embedding sum and the sampler: the Qwen3 layer stack that
 * AsteroidTTSInstruct.forward drives (modeling_asteroid.py:252-285: RMSNorm, q/k/v projection, per-head q/k RMSNorm +
 * RoPE, KV-cache append, causal GQA attention, o_proj + residual, RMSNorm, SwiGLU MLP + residual), the final norm and
 * the 8 lm_heads (:287-288). Same rounding points as the kernel chain (mtts_rmsnorm / mtts_gemm /
 * mtts_qknorm_rope_kvappend / mtts_gqa_attention), which it replaces launch for launch.
 * `layers` is a DEVICE array of num_layers descriptors; every matrix is row-major [out, in] bf16, wgu has gate/up rows
 * interleaved (2j = gate_j, 2j+1 = up_j), heads is [vpad, hidden]. x [B, hidden] bf16 holds the embedding sum (read
 * only); logits [B, ld_logits] bf16. workspace: mtts_decode_mega_workspace_bytes() bytes, 256-byte aligned, zeroed ONCE
 * at allocation and owned by this entry point afterwards (it carries the inter-CTA activation words and their tags).
 * The launch is cooperative (one CTA per SM); it fails with an error instead of hanging when the grid cannot be
 * co-resident. */
typedef struct mtts_lm_layer {
  const void* wqkv; const void* wo; const void* wgu; const void* wd;
  const void* ln1;  const void* ln2; const void* q_norm; const void* k_norm;
  void* k_pool;     void* v_pool;   /* [num_pages, num_kv_heads, page_size, head_dim] bf16 of this layer */
} mtts_lm_layer;

typedef struct mtts_decode_mega_args {
  const mtts_lm_layer* layers; int num_layers;
  int hidden, intermediate, num_q_heads, num_kv_heads, head_dim;
  const void* heads; int vpad;
  const void* final_norm;
  const float* inv_freq;          /* [head_dim/2] */
  const int* positions;           /* [B] position of the new row of each sequence */
  const int* block_table;         /* [B, max_pages] or NULL (contiguous cache) */
  int max_pages, page_size, num_pages;
  const void* x;
  void* logits; long long ld_logits;
  int B; int nsplit;              /* split-KV factor, 1..16; B * num_kv_heads * nsplit <= SM count */
  float eps;
  void* workspace; long long workspace_bytes;
  int* err_flag;                  /* set to 2 when a position falls outside the page table */
  long long* profile_cycles;      /* NULL, or [32 + 16*SMs] device int64 (profiling builds of the host only):
                                     SM cycles CTA 0 spent in each phase, then per-CTA globaltimer stamps of layer 5 */
} mtts_decode_mega_args;

int mtts_decode_mega_supported(int hidden, int intermediate, int num_q_heads, int num_kv_heads, int head_dim, int B);
long long mtts_decode_mega_workspace_bytes(int B, int nsplit);
int mtts_decode_mega(const mtts_decode_mega_args* args, void* stream);

/* im2col for Conv1d(K odd, pad=(K-1)/2, stride) on token-major x [B,T,Cin]: col[b,t',j*Cin+ci] = x[b,t'*stride+j-pad,ci]
 * (VocosBackbone.embed modules.py:1372; OmniAudioEncoder conv1/conv2 modules.py:238-240). */
int mtts_im2col(const float* x, float* col, int B, int T, int Cin, int K, int ld_col, int stride, void* stream);

/* Log-mel front end of XY_Tokenizer.encode (MelFeatureExtractor, nn/feature_extractor.py:78-104): torch.stft with a
 * centred, reflect-padded Hann window restated as framing (this kernel) + an exact-fp32 DFT GEMM (mtts_gemm_simt);
 * then |.|^2 (mtts_power_spectrum), the mel filter bank (mtts_gemm_simt) and log10 / per-item max-8 clamp / (x+4)/4
 * (mtts_logmel_finish, one CTA per item, in place on [B, per_item]). wav [B, L] fp32 rows of stride ld_wav. */
int mtts_stft_frames(const float* wav, long long ld_wav, const float* window, float* frames, int B, int T, int L, int n_fft,
                     int hop, void* stream);
int mtts_power_spectrum(const float* spec, long long lds, float* out, long long ldo, long long rows, int num_bins,
                        void* stream);
int mtts_logmel_finish(float* mel, int B, int per_item, void* stream);

/* ISTFTHead nonlinearity (modules.py:971-984): x [rows, 2F] = (log-mag | phase) -> spec [rows, lds] =
 * (Re_0..Re_{F-1} | Im_0..Im_{F-1} | 0...) with mag = min(exp(.), 100). */
int mtts_istft_spec(const float* x, long long ldx, float* spec, long long lds, long long rows, int num_bins,
                    void* stream);

/* ISTFT overlap-add, window-envelope normalisation and "same" trim (modules.py:759-792): frames [B,T,n_fft]
 * (already windowed: the inverse-DFT-times-window basis is applied by mtts_gemm) -> out [B, T*hop]. */
int mtts_istft_ola(const float* frames, const float* window, float* out, int B, int T, int n_fft, int hop, void* stream);

/* ISTFTHead.forward in one call (XY_Tokenizer/xy_tokenizer/nn/modules.py:939-988, called from Vocos.forward :1451-1479;
 * SURVEY.md 8b `mtts_istft_head`): wav [B, T*hop] = istft(x . head_w^T + head_b) with x [B*T, channels] fp32 (row stride
 * ldx), head_w [n_fft + 2, channels] (row stride ld_head_w), head_b [n_fft + 2], basis [n_fft, lds] the inverse real DFT
 * times the synthesis window over (Re | Im | 0-pad) columns, lds >= n_fft + 2 and a multiple of 4, window [n_fft].
 * = mtts_gemm(bias) -> mtts_istft_spec -> mtts_gemm -> mtts_istft_ola on `stream`; the three fp32 intermediates live in
 * `workspace` (mtts_istft_head_workspace_bytes() bytes, 256-byte aligned, caller-owned). */
size_t mtts_istft_head_workspace_bytes(int B, int T, int n_fft, long long lds);
int mtts_istft_head(const float* x, long long ldx, int channels, const float* head_w, long long ld_head_w,
                    const float* head_b, const float* basis, long long lds, const float* window, float* wav, int B, int T,
                    int n_fft, int hop, void* workspace, size_t workspace_bytes, void* stream);

/* x[r,:] += table[r % mod,:]  (sinusoidal positional embedding, modules.py:398-402,600-606). */
int mtts_add_rows_mod(float* x, const float* table, long long rows, int C, int mod, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTTS_H_ */
