// k_transition.cu — the split C-ABI transition call: one thread per env over the env-major state records (see
// transition.cuh for the logic and DESIGN.md for the measurements).  The per-env work is a few dozen dependent integer
// ops on ~200 B of state: at 8192 envs a launch is latency bound, at >= 1e6 envs it is an HBM streaming kernel.
#include "transition.cuh"

namespace cbs {

#ifndef CBS_TR_THREADS
#define CBS_TR_THREADS 128
#endif
#ifndef CBS_TR_MINB
#define CBS_TR_MINB 7
#endif
constexpr int TR_THREADS = CBS_TR_THREADS;
constexpr int TR_SC_SMEM = 64;   // scenario records kept in shared memory (larger sets are read through L2)

// Appends every thread's env to one of the SCHED_BINS decode cost bins (`bin` < 0: none) and to one of the OBS_CLASSES observe
// class lists (`cls` < 0: none) with ONE global atomic per list and CTA: positions inside the CTA come from shared-memory
// counters, the CTA's range from a single atomicAdd on the list's counter.
__device__ __forceinline__ void cta_append2(int bin, int cls, int b, int32_t* __restrict__ bin_cnt, int32_t* __restrict__ bin_list,
                                            int32_t* __restrict__ cls_cnt, int32_t* __restrict__ cls_list, int cap, int32_t* errflag,
                                            int* sh_cnt, int* sh_base) {
  constexpr int NL = SCHED_BINS + OBS_CLASSES;
  if (threadIdx.x < NL) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  int pos_bin = 0, pos_cls = 0;
  if (bin >= 0) pos_bin = atomicAdd(&sh_cnt[bin], 1);
  if (cls >= 0) pos_cls = atomicAdd(&sh_cnt[SCHED_BINS + cls], 1);
  __syncthreads();
  if (threadIdx.x < NL && sh_cnt[threadIdx.x]) {
    int32_t* ctr = threadIdx.x < SCHED_BINS ? &bin_cnt[threadIdx.x] : &cls_cnt[threadIdx.x - SCHED_BINS];
    sh_base[threadIdx.x] = atomicAdd(ctr, sh_cnt[threadIdx.x]);
  }
  __syncthreads();
  if (bin >= 0) {
    const int slot = sh_base[bin] + pos_bin;
    if (slot < cap) bin_list[(size_t)bin * cap + slot] = b;
  }
  if (cls >= 0) {
    const int slot = sh_base[SCHED_BINS + cls] + pos_cls;
    if (slot < cap) cls_list[(size_t)cls * cap + slot] = b; else atomicExch(errflag, 4);
  }
}

// One CTA = TR_THREADS consecutive envs, one thread per env.  Every thread requests its env's hot scalar sector, mask
// record, decoded action, distance, uniform and table-size estimate in one burst (TransitionIn::issue); measured with the
// transition logic knocked out, these per-thread record loads already stream at the HBM copy peak (27.6 us per 1M envs),
// so staging them through shared memory with 1-D bulk copies (tried: 99 us against 90 us) buys nothing — the kernel's
// time is the divergent per-env logic and its chain of L2 table look-ups.
template <bool DEF, bool REG>
__global__ void __launch_bounds__(TR_THREADS, DEF ? 1 : CBS_TR_MINB)
transition_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in, const double* __restrict__ dist_in,
                  const float* __restrict__ uniforms, int sched_out, float* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                  uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ outcome_out) {
  __shared__ int sh_cnt[SCHED_BINS + OBS_CLASSES], sh_base[SCHED_BINS + OBS_CLASSES];
  __shared__ int4 sh_sc[2 * TR_SC_SMEM];
  const int b = blockIdx.x * TR_THREADS + threadIdx.x;
  const bool live = b < P.B;
  TransitionIn<REG> in;
  int west = 0;
  if (live) {
    in.issue(P, S, b, uniforms);
    in.sl = reinterpret_cast<const int4*>(sel_in)[b];
    in.dist = dist_in ? dist_in[b] : 0.0;
    west = S.work_est[b];
  }
  // the scenario records (32 bytes each) go to shared memory while that burst is in flight: one look-up level less
  const bool sc_smem = T.num_scenarios <= TR_SC_SMEM;
  if (sc_smem) {
    for (int i = threadIdx.x; i < 2 * T.num_scenarios; i += TR_THREADS) sh_sc[i] = T.sc_pack[i];
    __syncthreads();
  }
  int cls = -1;
  if (live) {
    cls = transition_env<DEF, false, REG>(T, P, S, b, in, sc_smem ? sh_sc : T.sc_pack, uniforms != nullptr, sel_in != S.sel, sched_out,
                                          reward_out, done_out, trunc_out, outcome_out);
    in.store_hot(S, b);
  }
  // cost-binned env list for the next decode (longest tables first) and the observe kernel's class lists: the
  // two appends share their barriers
  cta_append2(live ? sched_bin(west) : -1, cls, b, S.bin_cnt + sched_out * (SCHED_BINS + 1),
              S.bin_list + (size_t)sched_out * SCHED_BINS * P.B, S.work_ctr + 4, S.worklist, P.B, S.errflag, sh_cnt, sh_base);
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
