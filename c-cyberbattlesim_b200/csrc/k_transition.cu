// k_transition.cu — the split C-ABI transition call: one thread per env over the env-major state records (see
// transition.cuh for the logic and DESIGN.md for the measurements).  The per-env work is a few dozen dependent integer
// ops on ~200 B of state: at 8192 envs a launch is latency bound, at >= 1e6 envs it is an HBM streaming kernel.
#include <cstdlib>

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
// class lists (`cls` < 0: none) with ONE global atomic per list and CTA — and without a CTA-wide barrier: every warp parks its
// entries in shared memory and leaves; the LAST warp of the CTA to arrive files all of them (positions inside the CTA from
// shared-memory counters, the CTA's range from a single atomicAdd on the list's counter).  With barriers here, a third of the
// kernel's warp samples were warps waiting for the CTA's slowest (most divergent) warp while holding their slots (ncu, round 2).
// `sh_cnt` (zeroed) and `sh_arrived` (0) must have been initialised before the kernel's first __syncthreads.
__device__ __forceinline__ void cta_append2(int bin, int cls, int32_t* __restrict__ bin_cnt, int32_t* __restrict__ bin_list,
                                            int32_t* __restrict__ cls_cnt, int32_t* __restrict__ cls_list, int cap, int32_t* errflag,
                                            int* sh_cnt, int* sh_base, int* sh_ent, int* sh_arrived, int* sh_env, int env) {
  constexpr int NL = SCHED_BINS + OBS_CLASSES, NW = TR_THREADS / 32;
  const int lane = threadIdx.x & 31;
  sh_ent[threadIdx.x] = (bin & 0xFF) | ((cls & 0xFF) << 8);     // -1 -> 0xFF = none
  sh_env[threadIdx.x] = env;
  __syncwarp();
  int last = 0;
  if (lane == 0) { __threadfence_block(); last = atomicAdd(sh_arrived, 1) == NW - 1; }
  last = __shfl_sync(0xFFFFFFFFu, last, 0);
  if (!last) return;
  __threadfence_block();
  int e[NW], pos_bin[NW], pos_cls[NW];
#pragma unroll
  for (int r = 0; r < NW; ++r) {
    e[r] = reinterpret_cast<volatile int*>(sh_ent)[r * 32 + lane];
    const int bn = e[r] & 0xFF, cl = (e[r] >> 8) & 0xFF;
    pos_bin[r] = bn != 0xFF ? atomicAdd(&sh_cnt[bn], 1) : 0;
    pos_cls[r] = cl != 0xFF ? atomicAdd(&sh_cnt[SCHED_BINS + cl], 1) : 0;
  }
  __syncwarp();
  if (lane < NL && sh_cnt[lane]) {
    int32_t* ctr = lane < SCHED_BINS ? &bin_cnt[lane] : &cls_cnt[lane - SCHED_BINS];
    sh_base[lane] = atomicAdd(ctr, sh_cnt[lane]);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NW; ++r) {
    const int bn = e[r] & 0xFF, cl = (e[r] >> 8) & 0xFF, b = reinterpret_cast<volatile int*>(sh_env)[r * 32 + lane];
    if (bn != 0xFF) {
      const int slot = sh_base[bn] + pos_bin[r];
      if (slot < cap) bin_list[(size_t)bn * cap + slot] = b;
    }
    if (cl != 0xFF) {
      const int slot = sh_base[SCHED_BINS + cl] + pos_cls[r];
      if (slot < cap) cls_list[(size_t)cl * cap + slot] = b; else atomicExch(errflag, 4);
    }
  }
}

// One CTA = TR_THREADS consecutive envs, one thread per env.  Every thread requests its env's hot scalar sector, mask
// record, decoded action, distance, uniform and table-size estimate in one burst (TransitionIn::issue); measured with the
// transition logic knocked out, these per-thread record loads already stream at the HBM copy peak (27.6 us per 1M envs),
// so staging them through shared memory with 1-D bulk copies (tried: 99 us against 90 us) buys nothing — the kernel's
// time is the divergent per-env logic and its chain of L2 table look-ups.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <bool DEF, bool REG>
__global__ void __launch_bounds__(TR_THREADS, DEF ? 1 : CBS_TR_MINB)
transition_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in, const double* __restrict__ dist_in,
                  const float* __restrict__ uniforms, int sched_out, float* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                  uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ outcome_out, int pf_dist) {
  __shared__ int sh_cnt[SCHED_BINS + OBS_CLASSES], sh_base[SCHED_BINS + OBS_CLASSES], sh_ent[TR_THREADS], sh_env[TR_THREADS], sh_arrived;
  __shared__ int4 sh_sc[2 * TR_SC_SMEM];
  if (threadIdx.x < SCHED_BINS + OBS_CLASSES) sh_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) sh_arrived = 0;
  const int b = blockIdx.x * TR_THREADS + threadIdx.x;
  const bool live = b < P.B;
  long long pf_b = -1;                       // env whose records are prefetched into L2 (`pf_dist` CTAs ahead)
  if (pf_dist > 0 && (long long)b + (long long)pf_dist * TR_THREADS < P.B) pf_b = (long long)b + (long long)pf_dist * TR_THREADS;
  TransitionIn<REG> in;
  int west = 0;
  if (live) {
    in.issue(P, S, b, uniforms);
    in.sl = reinterpret_cast<const int4*>(sel_in)[b];
    in.dist = dist_in ? dist_in[b] : 0.0;
    west = S.work_est[b];
  }
  if (pf_b >= 0) {
    // L2 prefetch of the records of the env `pf_dist` CTAs ahead (about one wave of resident CTAs): a prefetch holds no register
    // and no scoreboard entry, so the bytes in flight per SM are no longer capped by (resident threads x 160 B x the fraction of
    // a thread's life spent waiting for its burst) — which kept the launch at ~40 % of the copy peak although neither the
    // instruction issue nor the DRAM pipe was saturated.  The CTA that owns those envs later finds its burst in L2.
    const long long pb = pf_b;
    prefetch_l2(S.scal + (size_t)pb * 8);
    prefetch_l2(S.scal + ((size_t)P.B + pb) * 8);
    prefetch_l2(S.masks + (size_t)pb * P.mpitch);
    if (P.mpitch > 8) prefetch_l2(S.masks + (size_t)pb * P.mpitch + 8);
    if ((threadIdx.x & 1) == 0) prefetch_l2(sel_in + (size_t)pb * 4);            // 32-byte sectors: two envs' actions
    if ((threadIdx.x & 3) == 0 && dist_in) prefetch_l2(dist_in + pb);
    if ((threadIdx.x & 7) == 0) { prefetch_l2(S.work_est + pb); if (uniforms) prefetch_l2(uniforms + pb); }
  }
  // the scenario records (32 bytes each) go to shared memory while that burst is in flight: one look-up level less
  const bool sc_smem = T.num_scenarios <= TR_SC_SMEM;
  if (sc_smem)
    for (int i = threadIdx.x; i < 2 * T.num_scenarios; i += TR_THREADS) sh_sc[i] = T.sc_pack[i];
  __syncthreads();      // the kernel's only CTA-wide barrier
  int cls = -1;
  if (live) {
    cls = transition_env<DEF, false, REG>(T, P, S, b, in, sc_smem ? sh_sc : T.sc_pack, uniforms != nullptr, sel_in != S.sel, sched_out,
                                          reward_out, done_out, trunc_out, outcome_out);
    in.store_hot(S, b);
  }
  // cost-binned env list for the next decode (longest tables first) and the observe kernel's class lists
  cta_append2(live ? sched_bin(west) : -1, cls, S.bin_cnt + sched_out * (SCHED_BINS + 1),
              S.bin_list + (size_t)sched_out * SCHED_BINS * P.B, S.work_ctr + 4, S.worklist, P.B, S.errflag, sh_cnt, sh_base, sh_ent,
              &sh_arrived, sh_env, b);
}

// ------------------------------------------------------------------------------------------------------------------------
// transition_stream_kernel — the large-batch form of the same transition: persistent warps behind a TMA prefetch pipeline.
//
// Measured on the 1M-env launch (ncu, round 2): the thread-per-env kernel above reaches ~40 % of the copy peak although neither
// the DRAM pipe (33 %) nor the issue slots (41-50 %) are saturated — a thread spends a third of its life waiting for its own
// burst of record loads, so the bytes in flight per SM (resident threads x 160 B x that fraction) stay below what the HBM
// latency needs, and streaming time and logic time ADD UP (27 us + 41 us) instead of overlapping.  Here every warp owns a
// 5 KB shared-memory buffer and an mbarrier: lane 0 requests the records of the warp's NEXT group of 32 envs with 1-D bulk
// copies (cp.async.bulk -> SASS UBLKCP, completion by mbarrier transaction bytes) as soon as the current group has been moved
// from the buffer to registers, so the copy of group g + W runs under the logic of group g and the DRAM pipe never waits for a
// scoreboard.  No CTA-wide barrier anywhere: a warp's list entries (decode cost bins, observe classes) are parked in its own
// shared-memory slots and filed every STREAM_FLUSH groups with one global atomic per list.
// One-word planes, no defender (the record then is exactly 64 B of masks + one hot sector + half a counter sector).
// ------------------------------------------------------------------------------------------------------------------------
constexpr int STREAM_WARPS = 4;
constexpr int STREAM_FLUSH = 8;                       // groups between two filings of a warp's list entries
constexpr int SB_HOT = 0, SB_SEC1 = 1024, SB_MASK = 2048, SB_SEL = 4096, SB_DIST = 4608, SB_UNI = 4864, SB_WEST = 4992, SB_BYTES = 5120;

struct __align__(16) StreamWarp {
  unsigned char buf[SB_BYTES];
  int ent_env[32 * STREAM_FLUSH];
  unsigned short ent_code[32 * STREAM_FLUSH];
  int cnt[SCHED_BINS + OBS_CLASSES], base[SCHED_BINS + OBS_CLASSES];
  unsigned long long mbar;
};

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// bounded wait: a protocol bug must end the kernel with an error flag, never hang the GPU
__device__ __forceinline__ bool st_mbar_wait(uint32_t bar, uint32_t parity, int32_t* errflag) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  atomicExch(errflag, 9);
  return false;
}

// lane 0: request the records of envs [b0, b0 + 32) into the warp's buffer
__device__ __forceinline__ void stream_request(const Params& P, const State& S, StreamWarp& W, int b0, const int32_t* sel_in,
                                               const double* dist_in, const float* uniforms) {
  const uint32_t bar = st_smem_u32(&W.mbar), dst = st_smem_u32(W.buf);
  const uint32_t bytes = 1024u + 1024u + 2048u + 512u + (dist_in ? 256u : 0u) + (uniforms ? 128u : 0u) + 128u;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  st_bulk_g2s(dst + SB_HOT, S.scal + (size_t)b0 * 8, 1024u, bar);
  st_bulk_g2s(dst + SB_SEC1, S.scal + ((size_t)P.B + b0) * 8, 1024u, bar);
  st_bulk_g2s(dst + SB_MASK, S.masks + (size_t)b0 * 16, 2048u, bar);
  st_bulk_g2s(dst + SB_SEL, sel_in + (size_t)b0 * 4, 512u, bar);
  if (dist_in) st_bulk_g2s(dst + SB_DIST, dist_in + b0, 256u, bar);
  if (uniforms) st_bulk_g2s(dst + SB_UNI, uniforms + b0, 128u, bar);
  st_bulk_g2s(dst + SB_WEST, S.work_est + b0, 128u, bar);
}

// files the warp's parked list entries [0, n): one global atomic per non-empty list
__device__ __forceinline__ void stream_file(const Params& P, const State& S, StreamWarp& W, int n, int sched_out, int lane) {
  constexpr int NL = SCHED_BINS + OBS_CLASSES;
  if (lane < NL) W.cnt[lane] = 0;
  __syncwarp();
  int pos_bin[STREAM_FLUSH], pos_cls[STREAM_FLUSH];
#pragma unroll
  for (int r = 0; r < STREAM_FLUSH; ++r) {
    const int i = r * 32 + lane;
    pos_bin[r] = pos_cls[r] = 0;
    if (i < n) {
      const int bn = W.ent_code[i] & 0xFF, cl = W.ent_code[i] >> 8;
      if (bn != 0xFF) pos_bin[r] = atomicAdd(&W.cnt[bn], 1);
      if (cl != 0xFF) pos_cls[r] = atomicAdd(&W.cnt[SCHED_BINS + cl], 1);
    }
  }
  __syncwarp();
  if (lane < NL && W.cnt[lane]) {
    int32_t* ctr = lane < SCHED_BINS ? &S.bin_cnt[sched_out * (SCHED_BINS + 1) + lane] : &S.work_ctr[4 + lane - SCHED_BINS];
    W.base[lane] = atomicAdd(ctr, W.cnt[lane]);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < STREAM_FLUSH; ++r) {
    const int i = r * 32 + lane;
    if (i < n) {
      const int bn = W.ent_code[i] & 0xFF, cl = W.ent_code[i] >> 8, b = W.ent_env[i];
      if (bn != 0xFF) {
        const int slot = W.base[bn] + pos_bin[r];
        if (slot < P.B) S.bin_list[((size_t)sched_out * SCHED_BINS + bn) * P.B + slot] = b;
      }
      if (cl != 0xFF) {
        const int slot = W.base[SCHED_BINS + cl] + pos_cls[r];
        if (slot < P.B) S.worklist[(size_t)cl * P.B + slot] = b; else atomicExch(S.errflag, 4);
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(STREAM_WARPS * 32, CBS_TR_MINB)
transition_stream_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in, const double* __restrict__ dist_in,
                         const float* __restrict__ uniforms, int sched_out, float* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                         uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ outcome_out) {
  __shared__ StreamWarp sh_w[STREAM_WARPS];
  __shared__ int4 sh_sc[2 * TR_SC_SMEM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  StreamWarp& W = sh_w[warp];
  const int n_groups = (P.B + 31) / 32, total_warps = gridDim.x * STREAM_WARPS;
  int g = blockIdx.x * STREAM_WARPS + warp;
  const bool sc_smem = T.num_scenarios <= TR_SC_SMEM;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&W.mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the first group: a full group goes through the pipeline, the batch's ragged last group (fewer than 32 envs) is read directly
    if (g < n_groups && g * 32 + 32 <= P.B) stream_request(P, S, W, g * 32, sel_in, dist_in, uniforms);
  }
  if (sc_smem)
    for (int i = threadIdx.x; i < 2 * T.num_scenarios; i += STREAM_WARPS * 32) sh_sc[i] = T.sc_pack[i];
  __syncthreads();      // scenario records staged; the only CTA-wide barrier
  uint32_t parity = 0;
  int n_ent = 0;
  for (; g < n_groups; g += total_warps) {
    const int b = g * 32 + lane;
    const bool full = g * 32 + 32 <= P.B, live = b < P.B;
    TransitionIn<true> in;
    int west = 0;
    if (full) {
      st_mbar_wait(st_smem_u32(&W.mbar), parity, S.errflag);
      parity ^= 1u;
      const int4* hot = reinterpret_cast<const int4*>(W.buf + SB_HOT) + 2 * lane;
      in.h0 = hot[0]; in.h1 = hot[1];
      in.c0 = reinterpret_cast<const int4*>(W.buf + SB_SEC1)[2 * lane];
      const uint4* mk = reinterpret_cast<const uint4*>(W.buf + SB_MASK) + 4 * lane;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const uint4 v = mk[q]; in.M.m[4 * q] = v.x; in.M.m[4 * q + 1] = v.y; in.M.m[4 * q + 2] = v.z; in.M.m[4 * q + 3] = v.w; }
      in.M.rec = S.masks + (size_t)b * P.mpitch;
      in.M.dirty = 0u;
      in.sl = reinterpret_cast<const int4*>(W.buf + SB_SEL)[lane];
      in.dist = dist_in ? reinterpret_cast<const double*>(W.buf + SB_DIST)[lane] : 0.0;
      in.uniform = uniforms ? reinterpret_cast<const float*>(W.buf + SB_UNI)[lane] : 0.f;
      west = reinterpret_cast<const int*>(W.buf + SB_WEST)[lane];
      __syncwarp();           // every lane has taken its records out of the buffer ...
      const int gn = g + total_warps;
      if (lane == 0 && gn < n_groups && gn * 32 + 32 <= P.B) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // ... before the async proxy overwrites it
        stream_request(P, S, W, gn * 32, sel_in, dist_in, uniforms);
      }
    } else if (live) {
      in.issue(P, S, b, uniforms);
      in.sl = reinterpret_cast<const int4*>(sel_in)[b];
      in.dist = dist_in ? dist_in[b] : 0.0;
      west = S.work_est[b];
    }
    int cls = -1;
    if (live) {
      cls = transition_env<false, false, true>(T, P, S, b, in, sc_smem ? sh_sc : T.sc_pack, uniforms != nullptr, sel_in != S.sel, sched_out,
                                               reward_out, done_out, trunc_out, outcome_out);
      in.store_hot(S, b);
    }
    // park this group's list entries; file them every STREAM_FLUSH groups
    W.ent_env[n_ent + lane] = b;
    W.ent_code[n_ent + lane] = (unsigned short)((live ? sched_bin(west) : 0xFF) | ((cls & 0xFF) << 8));
    n_ent += 32;
    __syncwarp();
    if (n_ent == 32 * STREAM_FLUSH) { stream_file(P, S, W, n_ent, sched_out, lane); n_ent = 0; }
  }
  if (n_ent) stream_file(P, S, W, n_ent, sched_out, lane);
}

// K steps per launch (the "K-step persistent kernel over fixed action traces"): one thread per env keeps the env's
// mask record, hot scalar sector and list lengths in registers, applies K pre-decoded actions sel[k][b] (with their distances and
// uniforms, or Philox draws) and writes the records back once.  Per env-step only the action (16 B), distance (8 B), uniform
// (4 B) and the outputs (reward 4 B, done 1 B) move; the 144-byte state moves once per K steps.  No observe runs between the
// steps, so the visible graph / action table / work lists are NOT maintained and a finished env stays finished: this is the
// simulation core alone, for pre-decoded traces and for the transition roofline.  One-word planes, no defender.
__global__ void __launch_bounds__(TR_THREADS, CBS_TR_MINB)
transition_ksteps_kernel(Tables T, Params P, State S, const int32_t* __restrict__ sel_in, const double* __restrict__ dist_in,
                         const float* __restrict__ uniforms, int K, float* __restrict__ reward_out, uint8_t* __restrict__ done_out) {
  __shared__ int4 sh_sc[2 * TR_SC_SMEM];
  const int b = blockIdx.x * TR_THREADS + threadIdx.x;
  const bool live = b < P.B;
  TransitionIn<true> in;
  if (live) in.issue(P, S, b, nullptr);
  const bool sc_smem = T.num_scenarios <= TR_SC_SMEM;
  if (sc_smem) {
    for (int i = threadIdx.x; i < 2 * T.num_scenarios; i += TR_THREADS) sh_sc[i] = T.sc_pack[i];
    __syncthreads();
  }
  if (!live) return;
  const int4 c0_in = in.c0;
  // the next step's inputs are requested before the current step runs
  int4 sl_n = reinterpret_cast<const int4*>(sel_in)[b];
  double d_n = dist_in ? dist_in[b] : 0.0;
  float u_n = uniforms ? uniforms[b] : 0.f;
  for (int k = 0; k < K; ++k) {
    in.sl = sl_n; in.dist = d_n; in.uniform = u_n;
    if (k + 1 < K) {
      const size_t o = (size_t)(k + 1) * P.B + b;
      sl_n = reinterpret_cast<const int4*>(sel_in)[o];
      if (dist_in) d_n = dist_in[o];
      if (uniforms) u_n = uniforms[o];
    }
    const size_t o = (size_t)k * P.B;
    transition_env<false, false, true, true>(T, P, S, b, in, sc_smem ? sh_sc : T.sc_pack, uniforms != nullptr, false, 0,
                                             reward_out ? reward_out + o : nullptr, done_out ? done_out + o : nullptr, nullptr, nullptr);
  }
  in.M.close();
  int32_t* cnt = S.scal + ((size_t)P.B + b) * 8;
  if (in.c0.x != c0_in.x) cnt[S_N_DISC - 8] = in.c0.x;
  if (in.c0.y != c0_in.y) cnt[S_N_OWNED - 8] = in.c0.y;
  if (in.c0.z != c0_in.z) cnt[S_DISC_AMOUNT - 8] = in.c0.z;
  in.store_hot(S, b);
}

cudaError_t launch_transition_ksteps(const Tables& T, const Params& P, const State& S, const int32_t* sel, const double* dist,
                                     const float* uniforms, int K, float* reward, uint8_t* done, cudaStream_t stream) {
  const int grid = (P.B + TR_THREADS - 1) / TR_THREADS;
  transition_ksteps_kernel<<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, K, reward, done);
  return cudaGetLastError();
}

cudaError_t launch_transition(const Tables& T, const Params& P, const State& S, const int32_t* sel, const double* dist,
                              const float* uniforms, int sched_out, float* reward, uint8_t* done, uint8_t* trunc,
                              uint8_t* outcome, int num_sms, cudaStream_t stream) {
  const int grid = (P.B + TR_THREADS - 1) / TR_THREADS;
  // launches of several waves of CTAs prefetch the records of the next wave into L2 (CBS_TR_PF overrides the distance, in CTAs)
  static const int pf_env = getenv("CBS_TR_PF") ? atoi(getenv("CBS_TR_PF")) : -1;
  // CBS_TR_MODE=1 selects the persistent TMA-pipelined form for launches it applies to (measured equal to the prefetching
  // thread-per-env kernel, 77 vs 75 us per 1M envs: the launch is bound by the ALU pipe, not by how the records arrive)
  static const int mode_env = getenv("CBS_TR_MODE") ? atoi(getenv("CBS_TR_MODE")) : 0;
  const int wave = num_sms * CBS_TR_MINB;
  const int pf = pf_env >= 0 ? pf_env : (grid > 2 * wave ? wave : 0);
  if (P.defender) {
    transition_kernel<true, false><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome, 0);
  } else if (mode_env == 1 && P.words == 1 && P.mpitch == 16 && (((uintptr_t)sel | (uintptr_t)dist | (uintptr_t)uniforms) & 15u) == 0) {
    int sgrid = num_sms * CBS_TR_MINB;
    const int need = (P.B + 32 * STREAM_WARPS - 1) / (32 * STREAM_WARPS);
    if (sgrid > need) sgrid = need;
    transition_stream_kernel<<<sgrid, STREAM_WARPS * 32, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome);
  } else if (P.words == 1) {   // <= 32 nodes: mask record staged in registers
    transition_kernel<false, true><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome, pf);
  } else {
    transition_kernel<false, false><<<grid, TR_THREADS, 0, stream>>>(T, P, S, sel, dist, uniforms, sched_out, reward, done, trunc, outcome, pf);
  }
  return cudaGetLastError();
}

}  // namespace cbs
