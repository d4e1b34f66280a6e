// Error plumbing and device attribute cache for libmtts.
#include "common.cuh"
#include "mtts_internal.h"
#include <string.h>

static thread_local char g_err[1024] = "";

int mtts_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#include <stdlib.h>
bool mtts_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTTS_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

bool mtts_pdl_small_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTTS_PDL_SMALL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && mtts_pdl_enabled();
}

static int g_num_sms[64] = {0};

int mtts_num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_num_sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_num_sms[dev] = n;
  }
  return g_num_sms[dev];
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void mtts_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long mtts_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* mtts_last_error(void) { return g_err; }
extern "C" int mtts_version(void) { return MTTS_VERSION; }

extern "C" int mtts_init(void) {
  int dev = 0;
  MTTS_CUDA_CHECK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  MTTS_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MTTS_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return mtts_set_error(MTTS_ERR_UNSUPPORTED, "libmtts is built for sm_100a only; device %d is sm_%d%d", dev, major,
                          minor);
  (void)mtts_num_sms();
  static bool configured[64] = {false};
  if (dev < 64 && !configured[dev]) {
    int rc = 0;
    if ((rc = mtts_configure_gemm_tc())) return rc;
    if ((rc = mtts_configure_attention())) return rc;
    if ((rc = mtts_configure_rvq())) return rc;
    if ((rc = mtts_configure_codec())) return rc;
    if ((rc = mtts_configure_decode_mega())) return rc;
    if ((rc = mtts_configure_mha_tc5())) return rc;
    if ((rc = mtts_configure_prefill_tc5())) return rc;
    if ((rc = mtts_configure_sampler())) return rc;
    configured[dev] = true;
  }
  return MTTS_OK;
}
