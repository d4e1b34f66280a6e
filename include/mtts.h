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

/* GEMM epilogue flags (bitmask). Order of application:
 *   v = acc; +bias[n]; gelu; round to out dtype; *gamma[n]; residual[m,n] + v.
 * SWIGLU is exclusive: weight rows interleaved (2j = gate_j, 2j+1 = up_j), out[m, j] = silu(g) * u. */
#define MTTS_EPI_BIAS 1
#define MTTS_EPI_GELU 2
#define MTTS_EPI_GAMMA 4
#define MTTS_EPI_RESIDUAL 8
#define MTTS_EPI_SWIGLU 16

const char* mtts_last_error(void);
int mtts_version(void);
/* Caches per-device attributes for the current device. Safe to call repeatedly. */
int mtts_init(void);

/* ------------------------------------------------------------------------------------------------
 * Dense projections (tcgen05 / TMEM / TMA).   out[M,N] = epi(x[M,K] . w[N,K]^T)
 * Replaces aten::linear behind HF Qwen3Attention / Qwen3MLP (invoked from modeling_asteroid.py:273-284),
 * the 8 lm_heads (modeling_asteroid.py:412) and the codec nn.Linear / ConvTranspose1d layers
 * (XY_Tokenizer/xy_tokenizer/nn/modules.py:84-87,181-182,494-500,1111-1115,957).
 * in_dtype BF16: bf16 operands; F32: fp32 storage multiplied as TF32. fp32 accumulation either way.
 * bias/gamma are fp32 [N]; residual has the output dtype. Row strides in ELEMENTS.
 * workspace: mtts_gemm_workspace_bytes() bytes whose first 16 KiB are ZERO before the first use.
 * ---------------------------------------------------------------------------------------------- */
size_t mtts_gemm_workspace_bytes(int M, int N, int K, int in_dtype);
int mtts_gemm(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo, int M, int N,
              int K, int in_dtype, int out_dtype, int flags, const float* bias, const float* gamma,
              const void* residual, long long ldr, void* workspace, size_t workspace_bytes, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* MTTS_H_ */
