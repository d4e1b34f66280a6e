// Internal declarations shared by the translation units of libmtts.
#pragma once
#include "../../include/mtts.h"
#ifdef __CUDACC__
#include <cuda.h>
#endif

int mtts_gemm_tc_pick_bn(int M);
#ifdef __CUDACC__
// LM-heads GEMM whose epilogue reports, per batch row and 32-row quarter of the stacked head matrix, the best and the
// second-best bf16 logit as sortable keys [M][N/32][2] (gemm_tc.cu; used by mtts_heads8_sample, sampler.cu)
int mtts_gemm_heads_argmax(const void* x, long long ldx, const void* w, long long ldw, int M, int N, int K, int n_chan,
                           const int* chan_lo, const int* chan_hi, unsigned int* keys, cudaStream_t stream);
int mtts_get_tmap_2d(const void* ptr, long long rows, long long cols, long long ld, int box_rows, int elem_code, CUtensorMap* out);
#endif

// one-time per-device kernel attribute setup, called by mtts_init()
int mtts_configure_mha_tc5();
int mtts_configure_prefill_tc5();
int mtts_configure_sampler();
int mtts_configure_gemm_tc();
int mtts_configure_attention();
int mtts_configure_rvq();
int mtts_configure_codec();
int mtts_configure_decode_mega();
