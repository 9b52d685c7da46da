// tcgen05 weight-gradient kernel (placeholder until the MN-major operand path is validated on hardware):
// reports "not supported" so dp_conv_wgrad takes the CUDA-core split-K kernel.
#include "dp_common.cuh"
#include "conv_internal.cuh"

namespace dp {

bool tc_wgrad_supported(const dp_conv_desc*) { return false; }
size_t tc_wgrad_workspace(const dp_conv_desc*) { return 0; }
int tc_conv_wgrad(const dp_conv_desc*, const void*, const void*, float*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 wgrad: not available");
  return DP_ERR_UNSUPPORTED;
}

}  // namespace dp
