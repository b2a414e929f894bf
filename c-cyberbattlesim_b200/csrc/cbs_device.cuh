// cbs_device.cuh — small device helpers shared by the kernels.
#pragma once
#include "cbs_types.h"

namespace cbs {

__device__ __forceinline__ uint32_t ld_mask(const State& S, const Params& P, int plane, int w, int b) {
  return S.masks[(size_t)b * P.mpitch + plane * P.words + w];
}
__device__ __forceinline__ bool bit_of(const State& S, const Params& P, int plane, int node, int b) {
  return (ld_mask(S, P, plane, node >> 5, b) >> (node & 31)) & 1u;
}
__device__ __forceinline__ int32_t& scalar(const State& S, const Params& P, int plane, int b) {
  return S.scal[(size_t)b * SCAL_PITCH + plane];
}

__device__ __forceinline__ bool is_node_goal(const Params& P) { return P.goal >= GOAL_CONTROL_NODE; }
__device__ __forceinline__ int base_goal(const Params& P) { return is_node_goal(P) ? P.goal - GOAL_CONTROL_NODE : P.goal; }

// attacker_goal_reached (cyberbattle_env.py:467-514), on mask words.  `interest` = the env's interest node (*_node goals)
__device__ inline bool goal_reached(const State& S, const Params& P, int b, int interest) {
  if (P.goal == GOAL_CONTROL_NODE) return bit_of(S, P, M_OWNED, interest, b) && bit_of(S, P, M_PRIV_ROOT, interest, b);
  if (P.goal == GOAL_DISCOVERY_NODE)    // :493-508 (has_data is cleared by the collection, so "collected and exfiltrated" cannot hold with it)
    return bit_of(S, P, M_DISCOVERED, interest, b) && bit_of(S, P, M_VISIBLE, interest, b) &&
           (!bit_of(S, P, M_HAS_DATA, interest, b) ||
            (bit_of(S, P, M_COLLECTED, interest, b) && bit_of(S, P, M_EXFILTRATED, interest, b)));
  if (P.goal == GOAL_DISRUPTION_NODE) return bit_of(S, P, M_STOPPED, interest, b);
  const int starter = scalar(S, P, S_STARTER, b);
  int n_goal = 0, n_data = 0, n_pending = 0;
  for (int w = 0; w < P.words; ++w) {
    const uint32_t disc = ld_mask(S, P, M_DISCOVERED, w, b);
    const uint32_t not_starter = ((starter >> 5) == w) ? ~(1u << (starter & 31)) : 0xFFFFFFFFu;
    if (P.goal == GOAL_CONTROL) n_goal += __popc(ld_mask(S, P, M_OWNED, w, b) & ld_mask(S, P, M_PRIV_ROOT, w, b) & not_starter);
    else if (P.goal == GOAL_DISRUPTION) n_goal += __popc(disc & ld_mask(S, P, M_STOPPED, w, b) & not_starter);
    else {
      n_goal += __popc(disc & not_starter);
      n_data += __popc(disc & ld_mask(S, P, M_HAS_DATA, w, b));
      n_pending += __popc(disc & ld_mask(S, P, M_COLLECTED, w, b) & ~ld_mask(S, P, M_EXFILTRATED, w, b));
    }
  }
  if (P.goal == GOAL_CONTROL) return n_goal == scalar(S, P, S_OWNABLE, b);
  if (P.goal == GOAL_DISRUPTION) return n_goal == scalar(S, P, S_DISRUPTABLE, b);
  return n_goal == scalar(S, P, S_DISCOVERABLE, b) && n_data == 0 && n_pending == 0;
}

// cost bin of an env for the longest-first decode schedule
__device__ __forceinline__ int sched_bin(int rows) {
  const int b = 31 - __clz((rows >> 5) | 1);   // rows < 64 -> 0, < 128 -> 1, ...
  return b < SCHED_BINS - 1 ? b : SCHED_BINS - 1;
}
// The bins are double buffered: a decode reads (and finally clears) buffer `p` while the transitions of the same
// step fill buffer `p ^ 1` for the next decode.
__device__ __forceinline__ void sched_enqueue(const State& S, const Params& P, int b, int buf) {
  const int bin = sched_bin(S.work_est[b]);
  const int pos = atomicAdd(&S.bin_cnt[buf * (SCHED_BINS + 1) + bin], 1);
  if (pos < P.B) S.bin_list[((size_t)buf * SCHED_BINS + bin) * P.B + pos] = b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

}  // namespace cbs
