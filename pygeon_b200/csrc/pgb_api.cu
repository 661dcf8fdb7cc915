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
