// Philox4x32-10 counter-based generator (Salmon et al., SC'11).  Key = the handle's seed, counter =
// (global env index lo/hi, per-env step counter, stream id).  Stream ids: 0 success-rate draws
// (replaces random.random() at simulation/attacker_actions.py:190,409), 1 starter node, 2 scenario switch.
#pragma once
#include <cstdint>

namespace cbs {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ inline Philox4 philox4x32_10(uint64_t key, uint64_t ctr_lo, uint32_t ctr2, uint32_t ctr3) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr2, c3 = ctr3;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// 24-bit uniform in [0,1): exactly representable in float32 and float64
__host__ __device__ inline float philox_uniform(uint64_t key, uint64_t env, uint32_t step, uint32_t stream) {
  return (float)(philox4x32_10(key, env, step, stream).x >> 8) * (1.0f / 16777216.0f);
}

}  // namespace cbs
