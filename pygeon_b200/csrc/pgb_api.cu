// Error plumbing + version of the C ABI (include/pgb.h).
#include <stdarg.h>
#include <string.h>
#include "pgb_common.cuh"

static thread_local char g_err[512] = "";

void pgb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pgb_last_error(void) { return g_err; }
extern "C" int pgb_version(void) { return 100; }

/* Hint for the L2 fetch granularity (32, 64 or 128 bytes; cudaLimitMaxL2FetchGranularity).  Random
 * 32-byte sector gathers (K2 numeric) over-fetch from HBM at the default 64 bytes. */
extern "C" int pgb_set_l2_fetch_granularity(int bytes) {
  cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
  if (e != cudaSuccess) {
    pgb_set_error("cudaDeviceSetLimit(MaxL2FetchGranularity, %d): %s", bytes, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
extern "C" int pgb_get_l2_fetch_granularity(void) {
  size_t v = 0;
  cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
  return (int)v;
}
