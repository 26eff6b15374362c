// tcgen05 / TMEM / TMA implicit-GEMM convolution (placeholder until the kernel lands: reports "unsupported").
#include "common.cuh"
extern "C" {
int ich_conv_tc_supported(int, int, int, int, int, int, int, int, int) { return 0; }
int ich_conv_tc_fwd(const void*, int, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, int, void*) {
  ich_set_error("ich_conv_tc_fwd: not built");
  return 1;
}
int ich_conv_tc_wgrad_supported(int, int, int, int, int, int, int, int, int) { return 0; }
int ich_conv_tc_wgrad(const void*, int, const void*, int, float*, int, int, int, int, int, int, int, int, int, void*) {
  ich_set_error("ich_conv_tc_wgrad: not built");
  return 1;
}
}
