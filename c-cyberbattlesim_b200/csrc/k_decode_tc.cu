// k_decode_tc.cu — tcgen05 (TF32) contraction for the decode: placeholder until the kernel lands.
#include "cbs_types.h"

namespace cbs {
bool decode_gemm_tc_available() { return false; }
cudaError_t launch_decode_gemm_tc(const float*, const float*, float*, float*, int, int, int, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace cbs
