// placeholder: tcgen05 variant arrives in the next commit
#include "common.cuh"
namespace daam {
bool mma_supported(const LayerParams&) { return false; }
int launch_accumulate_mma(const LaunchParams&, const DeviceInfo&, cudaStream_t) {
  set_error("tcgen05 accumulate kernel not built");
  return DAAM_E_UNSUPPORTED;
}
}  // namespace daam
