// Internal declarations shared by the translation units of libmtts.
#pragma once
#include "../../include/mtts.h"

int mtts_gemm_tc_pick_bn(int M);
