// tcgen05 (TF32, TMEM accumulators) implementation of the skeleton-aware conv -- placeholder until bring-up.
#include "conv_common.cuh"
namespace hmvae {
bool conv_fprop_tc_supported(const hmvae_conv_plan*, int, int) { return false; }
int conv_fprop_tc(const hmvae_conv_plan*, const float*, const float*, const float*, float*, int, int, cudaStream_t) {
  return fail_arg("conv_fprop: tcgen05 path not built");
}
}  // namespace hmvae
