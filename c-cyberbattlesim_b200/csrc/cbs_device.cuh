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
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// scalar `plane` of env b: sector-major arrays, State::scal[sector][B][8]
__device__ __forceinline__ int32_t& scalar(const State& S, const Params& P, int plane, int b) {
  return S.scal[((size_t)(plane >> 3) * P.B + b) * 8 + (plane & 7)];
}

__device__ __forceinline__ bool is_node_goal(const Params& P) { return P.goal >= GOAL_CONTROL_NODE; }
__device__ __forceinline__ int base_goal(const Params& P) { return is_node_goal(P) ? P.goal - GOAL_CONTROL_NODE : P.goal; }

// rows that create_continuous_action_space never adds to the action table (compressed:532-547)
__device__ __forceinline__ bool row_filtered(const Params& P, int kind, int s, int t, int starter, int interest) {
  if ((s == t && kind == K_LATERAL) || kind == K_CREDACCESS) return true;                         // compressed:532
  if (kind != K_DOS) return false;
  if (P.remove_all && P.goal != GOAL_DISRUPTION && P.goal != GOAL_DISRUPTION_NODE) return true;    // :536-538
  if (P.remove_main && t == starter) return true;                                                 // :541-543
  if (P.remove_main && interest >= 0 && P.goal != GOAL_DISRUPTION_NODE && t == interest) return true;   // :545-547
  return false;
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
