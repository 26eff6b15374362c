// Error plumbing shared by every entry point of libich_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void ich_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ich_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ich_set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

extern "C" {
const char* ich_last_error(void) { return g_err; }
int ich_abi_version(void) { return 1; }
}
