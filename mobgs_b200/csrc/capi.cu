// Error plumbing + version for the C ABI (include/mobgs_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mobgs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MOBGS_ECUDA;
  }
  return MOBGS_OK;
}

}  // namespace mobgs

extern "C" const char* mobgs_version(void) { return "mobgs_b200 0.1 (sm_100a)"; }
extern "C" const char* mobgs_last_error(void) { return mobgs::g_err; }
