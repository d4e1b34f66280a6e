// Internal declarations shared by the translation units of libmtts.
#pragma once
#include "../../include/mtts.h"

int mtts_gemm_tc_pick_bn(int M);

// one-time per-device kernel attribute setup, called by mtts_init()
int mtts_configure_gemm_tc();
int mtts_configure_attention();
int mtts_configure_rvq();
int mtts_configure_codec();
int mtts_configure_decode_mega();
