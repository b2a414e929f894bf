// k_transition.cu — the split C-ABI transition call: one thread per env over the env-major state records (see
// transition.cuh for the logic and DESIGN.md for the measurements).  The per-env work is a few dozen dependent integer
// ops on ~200 B of state: at 8192 envs a launch is latency bound, at >= 1e6 envs it is an HBM streaming kernel.
#include "transition.cuh"

namespace cbs {

constexpr int TR_THREADS = 128;

// Appends every thread's env to one of NLISTS lists (`which` < 0: none) with ONE global atomic per list and CTA: positions
// inside the CTA come from shared-memory counters, the CTA's range from a single atomicAdd on the list's counter.
template <int NLISTS>
__device__ __forceinline__ void cta_append(int which, int b, int32_t* __restrict__ counters, int32_t* __restrict__ lists,
                                           int list_pitch, int cap, int32_t* errflag, int err, int* sh_cnt, int* sh_base) {
  if (threadIdx.x < NLISTS) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  int pos = 0;
  if (which >= 0) pos = atomicAdd(&sh_cnt[which], 1);
  __syncthreads();
  if (threadIdx.x < NLISTS) sh_base[threadIdx.x] = sh_cnt[threadIdx.x] ? atomicAdd(&counters[threadIdx.x], sh_cnt[threadIdx.x]) : 0;
  __syncthreads();
  if (which >= 0) {
    const int slot = sh_base[which] + pos;
    if (slot < cap) lists[(size_t)which * list_pitch + slot] = b;
    else if (errflag) atomicExch(errflag, err);
  }
}

template <bool DEF, bool REG>
__global__ void __launch_bounds__(TR_THREADS) transition_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in,
                                                                const double* __restrict__ dist_in,
                                                                const float* __restrict__ uniforms, int sched_out,
                                                                float* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                                                                uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ outcome_out) {
  __shared__ int sh_cnt[SCHED_BINS], sh_base[SCHED_BINS];
  const int b = blockIdx.x * TR_THREADS + threadIdx.x;
  const bool live = b < P.B;
  // cost-binned env list for the next decode (longest tables first)
  cta_append<SCHED_BINS>(live ? sched_bin(S.work_est[b]) : -1, b, S.bin_cnt + sched_out * (SCHED_BINS + 1),
                         S.bin_list + (size_t)sched_out * SCHED_BINS * P.B, P.B, P.B, nullptr, 0, sh_cnt, sh_base);
  int cls = -1;
  if (live)
    cls = transition_env<DEF, false, REG>(T, P, S, b, reinterpret_cast<const int4*>(sel_in)[b], dist_in ? dist_in[b] : 0.0, uniforms,
                                     sched_out, reward_out, done_out, trunc_out, outcome_out);
  // the observe kernel's three class lists
  cta_append<3>(cls, b, S.work_ctr + 4, S.worklist, P.B, P.B, S.errflag, 4, sh_cnt, sh_base);
}

cudaError_t launch_transition(const Tables& T, const Params& P, const State& S, const int32_t* sel, const double* dist,
                              const float* uniforms, int sched_out, float* reward, uint8_t* done, uint8_t* trunc,
                              uint8_t* outcome, cudaStream_t stream) {
  const int grid = (P.B + TR_THREADS - 1) / TR_THREADS;
  if (P.defender)
    transition_kernel<true, false><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  else if (P.words == 1)   // <= 32 nodes: mask record staged in registers
    transition_kernel<false, true><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  else
    transition_kernel<false, false><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  return cudaGetLastError();
}

}  // namespace cbs
