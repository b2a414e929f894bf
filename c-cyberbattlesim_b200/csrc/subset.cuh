// subset.cuh — sample_subset_samples: the sub-sampled, EXPLICIT action table (device functions).
//
// Reference: CyberBattleCompressedEnv.__balance_action_space_by_outcome (_env/cyberbattle_env_compressed.py:553-567), called at
// the end of every create_continuous_action_space when `sample_subset_samples` is set (:521-522; 100 in the reference's
// train_config.yaml:29).  The table's keys are grouped by outcome class in first-appearance order, a class with more than k rows
// keeps k of them, the table becomes the concatenation of the groups, and dropped rows never return (their pair stays in
// processed_pairs; only precise_action_space_positions re-adds them, at the end of the table).
//
// The reference draws the k rows with np.random.choice.  Here — and in the oracle, and in the reference run that recorded the
// golden traces (oracle/ref_bridge.py hands the same rule to the unmodified method) — row i of an over-full class gets the key
//     philox4x32_10(seed, global env index, c, identity_i).x ,   identity = 0x80000000 | source << 23 | target << 16 | kind << 12 | vuln
// with c = the env's lifetime count of balance calls, and the k smallest (key, table position) pairs stay
// (ccbs_b200/philox.py::subset_keep).
//
// Layout: with the option on, an env's table is at most 10 classes x k rows, so it is held as per-class lists of 32-bit entries
// (source | target << 7 | row within the pair's candidate list << 14 | insertion epoch << 22) instead of the implicit
// (processed pair x candidate rows of the target) product the full table uses.  A row's place in the reference's table order is
// (first-appearance rank of its class, insertion epoch, source position, target position, row) — all recoverable from the
// entry — so the lists themselves are unordered.  One warp per env; the scratch (histogram, candidates, counters, the pairs this
// build touches) lives in the warp's shared-memory buffer that the encoder's GCN projection no longer needs.
#pragma once
#include "cbs_device.cuh"
#include "philox.cuh"

namespace cbs {

constexpr uint32_t SUBSET_STREAM = 0x80000000u;
constexpr int SUB_CAND = 128;   // candidates sorted directly once a radix bucket is this small

__device__ __forceinline__ uint32_t sub_entry(int s, int t, int rip, int epoch) {
  return (uint32_t)s | ((uint32_t)t << 7) | ((uint32_t)rip << 14) | ((uint32_t)epoch << 22);
}
__device__ __forceinline__ uint32_t sub_key32(const Params& P, uint64_t genv, uint32_t call, int s, int t, int kind, int ulocal) {
  const uint32_t id = SUBSET_STREAM | ((uint32_t)s << 23) | ((uint32_t)t << 16) | ((uint32_t)kind << 12) | (uint32_t)ulocal;
  return philox4x32_10(P.seed, genv, call, id).x;
}
// first-appearance rank of an outcome class in the env's table (15 = the class has no row yet)
__device__ __forceinline__ int sub_rank(const int32_t* __restrict__ meta, int kind) {
  return (int)((((uint32_t)meta[11 + (kind >> 3)]) >> (4 * (kind & 7))) & 15u);
}

// the warp's scratch, carved out of an 8 KB shared-memory buffer
struct SubScratch {
  uint32_t* hist;               // [256]
  unsigned long long* cand;     // [SUB_CAND]
  unsigned long long* thr;      // [SUB_CLASSES] largest (key, insertion key) that stays, ~0 = everything stays
  uint32_t* cnt;                // [SUB_CLASSES] rows in the list
  uint32_t* nnew;               // [SUB_CLASSES] rows this build adds
  uint32_t* first;              // [SUB_CLASSES] smallest insertion key among them
  uint32_t* ncand;              // [1]
  uint8_t* opos;                // [MAX_NODES] node -> position in owned_order
  int* ch_r0;                   // [32] staged chunk of touched pairs: first candidate row of the pair
  int* ch_pre;                  // [33] exclusive prefix of the chunk's row counts
  uint32_t* ch_meta;            // [32] source | target << 7 | insertion position of the source << 14 | target position << 22 | refreshed << 30
  uint16_t* newp;               // pairs this build touches: target position | source position << 7 | refreshed-not-new << 15
  __device__ __forceinline__ void carve(unsigned char* base, uint16_t* newp_global) {
    hist = reinterpret_cast<uint32_t*>(base);
    cand = reinterpret_cast<unsigned long long*>(base + 1024);
    thr = reinterpret_cast<unsigned long long*>(base + 2048);
    cnt = reinterpret_cast<uint32_t*>(base + 2128);
    nnew = cnt + SUB_CLASSES;
    first = nnew + SUB_CLASSES;
    ncand = first + SUB_CLASSES;
    opos = base + 2304;
    ch_r0 = reinterpret_cast<int*>(base + 2432);
    ch_pre = ch_r0 + 32;
    ch_meta = reinterpret_cast<uint32_t*>(ch_pre + 33);
    newp = newp_global ? newp_global : reinterpret_cast<uint16_t*>(base + 2832);
  }
};
constexpr int SUB_NEWP_SMEM = (8192 - 2832) / 2;   // 2680 >= 32 x 32 pairs

struct SubCtx {
  const Tables& T;
  const Params& P;
  const State& S;
  SubScratch sc;
  uint32_t* lists;          // this env's [SUB_CLASSES][K]
  uint32_t* alive;          // this env's [ncap*ncap][8] or nullptr
  const uint8_t* oorder;    // source list of this build (owned_order, or owned_raw under a defender)
  const uint8_t* dorder;    // discovered order
  const uint8_t* dpos;      // node -> position in the discovered order
  const uint8_t* popos;     // pair_opos of this env (defender) or nullptr
  uint64_t genv;
  uint32_t call;
  int b, lane, slot, n_newp, node_off, starter, interest, K;
};

__device__ __forceinline__ bool sub_alive_bit(const SubCtx& C, int s, int t, int rip) {
  return (C.alive[((size_t)s * C.P.ncap + t) * (SUB_MAX_ROWS_PER_PAIR / 32) + (rip >> 5)] >> (rip & 31)) & 1u;
}

// g(kind, s, t, rip, r, opk, dp) for every row the pairs of this build add (all classes), the rows flattened over the lanes: the
// touched pairs are staged 32 at a time (lane per pair: row range and identity), then lane j takes row j of the chunk, so that a
// trip's 32 template loads are independent and in flight together (a lane walking the rows of "its" pair made every row a
// dependent L2 round trip).  Rows that are filtered out of the table, or (refreshed pairs) still alive, are skipped.
template <class G>
__device__ __forceinline__ void sub_for_each_new(const SubCtx& C, G&& g) {
  const int ncap = C.P.ncap;
  for (int base = 0; base < C.n_newp; base += 32) {
    const int pi = base + C.lane;
    int cnt = 0;
    __syncwarp();
    if (pi < C.n_newp) {
      const uint32_t pr = C.sc.newp[pi];
      const int dp = pr & 127, op = (pr >> 7) & 255;
      const int s = C.oorder[op], t = C.dorder[dp];
      const int gn = C.node_off + t;
      const int r0 = C.T.nd_row_off[2 * gn + (s == t ? 0 : 1)], r1 = C.T.nd_row_off[2 * gn + 2];
      const int opk = C.popos ? (int)C.popos[s * ncap + t] : op;
      C.sc.ch_r0[C.lane] = r0;
      C.sc.ch_meta[C.lane] = (uint32_t)s | ((uint32_t)t << 7) | ((uint32_t)opk << 14) | ((uint32_t)dp << 22) | ((pr >> 15) << 30);
      cnt = r1 - r0;
    }
    int run = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, run, o);
      if (C.lane >= o) run += v;
    }
    C.sc.ch_pre[C.lane + 1] = run;
    if (C.lane == 0) C.sc.ch_pre[0] = 0;
    __syncwarp();
    const int total = C.sc.ch_pre[32];
    int q = 0;
    for (int j = C.lane; j < total; j += 32) {
      while (C.sc.ch_pre[q + 1] <= j) ++q;
      const int rip = j - C.sc.ch_pre[q];
      const int r = C.sc.ch_r0[q] + rip;
      const uint32_t m = C.sc.ch_meta[q];
      const int s = m & 127, t = (m >> 7) & 127, opk = (m >> 14) & 255, dp = (m >> 22) & 255;
      const int kind = (C.T.row_packed[r] >> 20) & 15;
      if (row_filtered(C.P, kind, s, t, C.starter, C.interest)) continue;
      if ((m >> 30) && sub_alive_bit(C, s, t, rip)) continue;      // still in the table: overwritten in place, not re-added
      g(kind, s, t, rip, r, opk, dp);
    }
  }
  __syncwarp();
}

// f(composite key, entry) for every row of class c: the rows in its list, then the rows the pairs of this build add.
// Lanes work independently (divergent): f may only use atomics on shared memory.
template <class F>
__device__ __forceinline__ void sub_for_each(const SubCtx& C, int c, int n_old, F&& f) {
  const int ncap = C.P.ncap;
  for (int i = C.lane; i < n_old; i += 32) {
    const uint32_t e = C.lists[c * C.K + i];
    const int s = e & 127, t = (e >> 7) & 127, rip = (e >> 14) & 255, epoch = (int)(e >> 22);
    const int g = C.node_off + t;
    const int r = C.T.nd_row_off[2 * g + (s == t ? 0 : 1)] + rip;
    const int opk = C.popos ? (int)C.popos[s * ncap + t] : (int)C.sc.opos[s];
    const uint32_t ins = ((uint32_t)epoch << 24) | ((uint32_t)opk << 16) | ((uint32_t)C.dpos[t] << 8) | (uint32_t)rip;
    const uint32_t k32 = sub_key32(C.P, C.genv, C.call, s, t, c, C.T.row_ulocal[r]);
    f(((unsigned long long)k32 << 32) | ins, e);
  }
  sub_for_each_new(C, [&](int kind, int s, int t, int rip, int r, int opk, int dp) {
    if (kind != c) return;
    const uint32_t ins = ((uint32_t)C.slot << 24) | ((uint32_t)opk << 16) | ((uint32_t)dp << 8) | (uint32_t)rip;
    const uint32_t k32 = sub_key32(C.P, C.genv, C.call, s, t, c, C.T.row_ulocal[r]);
    f(((unsigned long long)k32 << 32) | ins, sub_entry(s, t, rip, C.slot));
  });
}

// the `need`-th smallest composite key among the rows of class c (radix select, 8 bits per level; a bucket of at most SUB_CAND
// rows is ranked directly)
static __device__ unsigned long long sub_select(const SubCtx& C, int c, int n_old, int need, int n_total) {
  // Fast path: the keys are uniform 32-bit draws, so the need-th smallest of n_total rows lies near need / n_total * 2^32.  One
  // enumeration collects every row below 1.4x that estimate (+ slack) into the 256-entry buffer (histogram and candidate
  // areas together); if at least `need` and at most 256 rows qualified, the need-th smallest of them is the answer.
  if (need <= 160) {
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(C.sc.hist);      // hist (1 KB) + cand (1 KB) are contiguous
    const double est = 1.4 * (double)need / (double)n_total + 24.0 / (double)n_total;
    const unsigned long long cut = est >= 1.0 ? ~0ull : (unsigned long long)(est * 18446744073709551616.0);
    if (C.lane == 0) *C.sc.ncand = 0u;
    __syncwarp();
    sub_for_each(C, c, n_old, [&](unsigned long long comp, uint32_t) {
      if (comp <= cut) {
        const uint32_t at = atomicAdd(C.sc.ncand, 1u);
        if (at < 256u) buf[at] = comp;
      }
    });
    __syncwarp();
    const int m = (int)*C.sc.ncand;
    if (m >= need && m <= 256) {
      unsigned long long res = 0ull;
      for (int i = C.lane; i < m; i += 32) {
        const unsigned long long x = buf[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += buf[j] < x;
        if (rank == need - 1) res = x;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, res, o);
        res = v > res ? v : res;
      }
      __syncwarp();
      return res;
    }
    __syncwarp();
  }
  unsigned long long prefix = 0ull;
  int nbits = 0, bucket = 0;
  for (;;) {
    for (int i = C.lane; i < 256; i += 32) C.sc.hist[i] = 0u;
    __syncwarp();
    sub_for_each(C, c, n_old, [&](unsigned long long comp, uint32_t) {
      if (nbits == 0 || (comp >> (64 - nbits)) == (prefix >> (64 - nbits))) atomicAdd(&C.sc.hist[(comp >> (56 - nbits)) & 255u], 1u);
    });
    __syncwarp();
    // the bucket that holds the need-th smallest: lane l owns buckets [8 l, 8 l + 8)
    uint32_t h[8], mine = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) { h[k] = C.sc.hist[8 * C.lane + k]; mine += h[k]; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (C.lane >= o) incl += v;
    }
    const unsigned owner_mask = __ballot_sync(0xFFFFFFFFu, incl >= (uint32_t)need);
    const int owner = __ffs(owner_mask) - 1;
    int bsel = 0, below = 0, inb = 0;
    if (C.lane == owner) {
      uint32_t run = incl - mine;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (run + h[k] >= (uint32_t)need && inb == 0) { bsel = 8 * C.lane + k; below = (int)run; inb = (int)h[k]; }
        run += h[k];
      }
    }
    bsel = __shfl_sync(0xFFFFFFFFu, bsel, owner);
    below = __shfl_sync(0xFFFFFFFFu, below, owner);
    inb = __shfl_sync(0xFFFFFFFFu, inb, owner);
    need -= below;
    prefix |= (unsigned long long)bsel << (56 - nbits);
    nbits += 8;
    bucket = inb;
    if (bucket <= SUB_CAND || nbits == 64) break;
  }
  if (bucket > SUB_CAND) {          // cannot happen: composite keys are unique
    if (C.lane == 0) atomicExch(C.S.errflag, 6);
    return prefix;
  }
  if (C.lane == 0) *C.sc.ncand = 0u;
  __syncwarp();
  sub_for_each(C, c, n_old, [&](unsigned long long comp, uint32_t) {
    if (nbits == 64 ? comp == prefix : (comp >> (64 - nbits)) == (prefix >> (64 - nbits))) {
      const uint32_t at = atomicAdd(C.sc.ncand, 1u);
      if (at < (uint32_t)SUB_CAND) C.sc.cand[at] = comp;
    }
  });
  __syncwarp();
  const int m = (int)*C.sc.ncand;
  unsigned long long res = 0ull;
  for (int i = C.lane; i < m; i += 32) {
    const unsigned long long x = C.sc.cand[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += C.sc.cand[j] < x;
    if (rank == need - 1) res = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, res, o);
    res = v > res ? v : res;
  }
  __syncwarp();
  return res;
}

// Adds the rows of the pairs in sc.newp[0, n_newp) to the lists and balances every class (one warp).  Called by every
// create_continuous_action_space (reset included, and also when no pair is new: the balance counter still advances).
// `oorder` / `dorder` / `dpos`: the source list of this build, the discovered order and its inverse.
// `skipped`: table builds the env skipped since the last one that ran (FL_PENDING_SHIFT): each was a balance call of the reference.
static __device__ void subset_update(const Tables& T, const Params& P, const State& S, SubScratch sc, int b, int lane, int slot, int n_newp,
                              const uint8_t* oorder, int n_sources, const uint8_t* dorder, const uint8_t* dpos, int skipped) {
  int32_t* meta = S.sub_meta + (size_t)b * SUB_META;
  if (skipped) {
    if (lane == 0) meta[13] += skipped;
    __syncwarp();
  }
  const int sc_id = scalar(S, P, S_SCENARIO, b);
  SubCtx C{T, P, S, sc,
           S.sub_rows + (size_t)b * SUB_CLASSES * P.subset_k,
           P.precise_positions ? S.sub_alive + (size_t)b * P.ncap * P.ncap * (SUB_MAX_ROWS_PER_PAIR / 32) : nullptr,
           oorder, dorder, dpos,
           P.defender ? S.pair_opos + (size_t)b * P.ncap * P.ncap : nullptr,
           (uint64_t)(P.global_env_offset + b), (uint32_t)meta[13],
           b, lane, slot, n_newp, scalar(S, P, S_NODE_OFF, b), scalar(S, P, S_STARTER, b),
           is_node_goal(P) ? T.sc_interest[sc_id] : -1, P.subset_k};
  const int K = P.subset_k;
  if (lane < SUB_CLASSES) { sc.cnt[lane] = (uint32_t)meta[lane]; sc.nnew[lane] = 0u; sc.first[lane] = 0xFFFFFFFFu; sc.thr[lane] = ~0ull; }
  if (!P.defender)
    for (int i = lane; i < n_sources; i += 32) sc.opos[oorder[i]] = (uint8_t)i;
  __syncwarp();
  // ---- pass A: rows per class this build adds, and where each class first appears ----
  sub_for_each_new(C, [&](int kind, int, int, int rip, int, int opk, int dp) {
    atomicAdd(&sc.nnew[kind], 1u);
    atomicMin(&sc.first[kind], ((uint32_t)slot << 24) | ((uint32_t)opk << 16) | ((uint32_t)dp << 8) | (uint32_t)rip);
  });
  __syncwarp();
  // ---- classes that enter the table take the next ranks, in the order of their first rows ----
  if (lane == 0) {
    uint32_t lo = (uint32_t)meta[11], hi = (uint32_t)meta[12];
    int ranked = meta[10];
    for (;;) {
      int pick = -1;
      uint32_t best = 0xFFFFFFFFu;
      for (int k = 0; k < SUB_CLASSES; ++k) {
        const uint32_t rk = ((k < 8 ? lo : hi) >> (4 * (k & 7))) & 15u;
        if (rk == 15u && sc.nnew[k] > 0u && sc.first[k] < best) { best = sc.first[k]; pick = k; }
      }
      if (pick < 0) break;
      uint32_t& w = pick < 8 ? lo : hi;
      w = (w & ~(15u << (4 * (pick & 7)))) | ((uint32_t)ranked << (4 * (pick & 7)));
      ++ranked;
    }
    meta[10] = ranked; meta[11] = (int32_t)lo; meta[12] = (int32_t)hi;
  }
  __syncwarp();
  // ---- over-full classes: the k-th smallest (key, insertion key) is the largest that stays ----
  for (int c = 0; c < SUB_CLASSES; ++c) {
    const int n_old = (int)sc.cnt[c], n_new = (int)sc.nnew[c];
    if (n_new == 0 || n_old + n_new <= K) continue;
    const unsigned long long thr = sub_select(C, c, n_old, K, n_old + n_new);
    // compact the list in place
    int kept = 0;
    for (int base = 0; base < n_old; base += 32) {
      const int i = base + lane;
      uint32_t e = 0u;
      bool keep = false;
      if (i < n_old) {
        e = C.lists[c * K + i];
        const int s = e & 127, t = (e >> 7) & 127, rip = (e >> 14) & 255, epoch = (int)(e >> 22);
        const int g = C.node_off + t;
        const int r = T.nd_row_off[2 * g + (s == t ? 0 : 1)] + rip;
        const int opk = C.popos ? (int)C.popos[s * P.ncap + t] : (int)sc.opos[s];
        const uint32_t ins = ((uint32_t)epoch << 24) | ((uint32_t)opk << 16) | ((uint32_t)dpos[t] << 8) | (uint32_t)rip;
        const uint32_t k32 = sub_key32(P, C.genv, C.call, s, t, c, T.row_ulocal[r]);
        keep = ((((unsigned long long)k32) << 32) | ins) <= thr;
        if (!keep && C.alive)
          atomicAnd(&C.alive[((size_t)s * P.ncap + t) * (SUB_MAX_ROWS_PER_PAIR / 32) + (rip >> 5)], ~(1u << (rip & 31)));
      }
      const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
      __syncwarp();
      if (keep) C.lists[c * K + kept + __popc(km & ((1u << lane) - 1u))] = e;
      kept += __popc(km);
      __syncwarp();
    }
    if (lane == 0) { sc.cnt[c] = (uint32_t)kept; sc.thr[c] = thr; }
    __syncwarp();
  }
  // ---- pass B: the surviving new rows join their lists ----
  sub_for_each_new(C, [&](int kind, int s, int t, int rip, int r, int opk, int dp) {
    const unsigned long long thr = sc.thr[kind];
    if (thr != ~0ull) {
      const uint32_t ins = ((uint32_t)slot << 24) | ((uint32_t)opk << 16) | ((uint32_t)dp << 8) | (uint32_t)rip;
      const uint32_t k32 = sub_key32(P, C.genv, C.call, s, t, kind, T.row_ulocal[r]);
      if (((((unsigned long long)k32) << 32) | ins) > thr) return;
    }
    const uint32_t at = atomicAdd(&sc.cnt[kind], 1u);
    if (at < (uint32_t)K) {
      C.lists[kind * K + at] = sub_entry(s, t, rip, slot);
      if (C.alive) atomicOr(&C.alive[((size_t)s * P.ncap + t) * (SUB_MAX_ROWS_PER_PAIR / 32) + (rip >> 5)], 1u << (rip & 31));
    } else {
      atomicExch(S.errflag, 6);
    }
  });
  __syncwarp();
  uint32_t total = lane < SUB_CLASSES ? (sc.cnt[lane] < (uint32_t)K ? sc.cnt[lane] : (uint32_t)K) : 0u;
  if (lane < SUB_CLASSES) meta[lane] = (int32_t)total;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
  if (lane == 0) {
    meta[13] = (int32_t)(C.call + 1u);
    S.work_est[b] = (int32_t)total;
  }
  __syncwarp();
}

}  // namespace cbs
