// Error plumbing shared by all C-ABI entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace univs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return UNIVS_E_LAUNCH;
  }
  return UNIVS_OK;
}

}  // namespace univs

extern "C" const char* univs_b200_last_error(void) { return univs::g_err; }
extern "C" int univs_b200_abi_version(void) { return 1; }
