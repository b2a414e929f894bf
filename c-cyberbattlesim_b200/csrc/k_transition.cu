// k_transition.cu — the split C-ABI transition call: one thread per env over the SoA planes.  The per-env work is a
// few dozen dependent integer ops on ~200 B of state, so the kernel is bound by load latency (see transition.cuh for
// the logic and DESIGN.md for the measurements).
#include "transition.cuh"

namespace cbs {

template <bool DEF>
__global__ void __launch_bounds__(128) transition_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in,
                                                         const double* __restrict__ dist_in,
                                                         const float* __restrict__ uniforms, int sched_out,
                                                         float* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                                                         uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ outcome_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  transition_env<DEF>(T, P, S, b, reinterpret_cast<const int4*>(sel_in)[b], dist_in ? dist_in[b] : 0.0, uniforms, sched_out,
                 reward_out, done_out, trunc_out, outcome_out);
}

cudaError_t launch_transition(const Tables& T, const Params& P, const State& S, const int32_t* sel, const double* dist,
                              const float* uniforms, int sched_out, float* reward, uint8_t* done, uint8_t* trunc,
                              uint8_t* outcome, cudaStream_t stream) {
  if (P.defender)
    transition_kernel<true><<<(P.B + 127) / 128, 128, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  else
    transition_kernel<false><<<(P.B + 127) / 128, 128, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  return cudaGetLastError();
}

}  // namespace cbs
