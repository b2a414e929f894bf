// transition.cuh — state transition, reward and termination for ONE env (device function).
//
// Follows CyberBattleEnv.step_attacker_env (_env/cyberbattle_env.py:299-394) and
// AttackerAgentActions.exploit_{remote,local}_vulnerability (simulation/attacker_actions.py:92-547) on
// per-env bitmask planes and the flattened scenario tables.  Called by transition_kernel (one thread per env:
// the split C-ABI call) and by lane 0 of decode_select's warp right after the decode (the fused cbs_step path,
// where its chain of dependent loads hides among the other warps' table scans).
#pragma once
#include "cbs_device.cuh"
#include "philox.cuh"
#include <type_traits>

namespace cbs {

namespace {

// the env's mask record, accessed in place (any number of words per plane)
struct EnvBits {
  uint32_t* masks;   // this env's mask record
  int words;
  __device__ void open(uint32_t* rec, int w) { masks = rec; words = w; }
  __device__ void close() {}
  __device__ uint32_t word(int plane, int w) const { return masks[plane * words + w]; }
  __device__ bool get(int plane, int node) const { return (word(plane, node >> 5) >> (node & 31)) & 1u; }
  __device__ void set(int plane, int node) { masks[plane * words + (node >> 5)] |= (1u << (node & 31)); }
  __device__ void clr(int plane, int node) { masks[plane * words + (node >> 5)] &= ~(1u << (node & 31)); }
  __device__ void set_if(int plane, int node, bool c) { if (c) set(plane, node); }
  __device__ void clr_if(int plane, int node, bool c) { if (c) clr(plane, node); }
  __device__ void or_word(int plane, uint32_t bits) { masks[plane * words] |= bits; }
};

// Scenarios of <= 32 nodes (one word per plane): the whole 64-byte record is fetched with four 128-bit loads when the
// env is opened, lives in registers (every plane index below is a compile-time constant), and only the words that
// changed are written back.  All the state a transition needs is then requested in one burst instead of one dependent
// round trip per bit test.
struct EnvBitsReg {
  uint32_t m[16];
  uint32_t dirty;
  uint32_t* rec;
  __device__ __forceinline__ void open(uint32_t* r, int) {
    rec = r; dirty = 0u;
    const uint4* q = reinterpret_cast<const uint4*>(r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const uint4 v = q[i]; m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w; }
  }
  __device__ __forceinline__ void close() {
#pragma unroll
    for (int p = 0; p < N_MASKS; ++p) if ((dirty >> p) & 1u) rec[p] = m[p];
  }
  __device__ __forceinline__ uint32_t word(int plane, int) const { return m[plane]; }
  __device__ __forceinline__ bool get(int plane, int node) const { return (m[plane] >> (node & 31)) & 1u; }
  __device__ __forceinline__ void set(int plane, int node) {
    const uint32_t v = m[plane] | (1u << (node & 31));
    if (v != m[plane]) { m[plane] = v; dirty |= 1u << plane; }
  }
  __device__ __forceinline__ void clr(int plane, int node) {
    const uint32_t v = m[plane] & ~(1u << (node & 31));
    if (v != m[plane]) { m[plane] = v; dirty |= 1u << plane; }
  }
  // predicated forms: straight-line selects, no branch (the caller's condition implies that the bit really changes)
  __device__ __forceinline__ void set_if(int plane, int node, bool c) {
    m[plane] |= c ? (1u << (node & 31)) : 0u;
    dirty |= c ? (1u << plane) : 0u;
  }
  __device__ __forceinline__ void clr_if(int plane, int node, bool c) {
    m[plane] &= c ? ~(1u << (node & 31)) : 0xFFFFFFFFu;
    dirty |= c ? (1u << plane) : 0u;
  }
  __device__ __forceinline__ void or_word(int plane, uint32_t bits) {
    m[plane] |= bits;
    dirty |= bits ? (1u << plane) : 0u;
  }
};
static_assert(N_MASKS <= 16, "EnvBitsReg holds one 64-byte record");

// Philox draw j of a stream family: component (j & 3) of stream `base + (j >> 2)`
__device__ __forceinline__ uint32_t philox_word(uint64_t key, uint64_t env, uint32_t step, uint32_t base, int j) {
  const Philox4 r = philox4x32_10(key, env, step, base + (uint32_t)(j >> 2));
  const int c = j & 3;
  return c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
}

// ---- static defender step (cyberbattle_env.py:416-430) with ScanAndReimageCompromisedMachines (static_defender.py:45-60)
//      and StaticDefenderAgentActions (static_defender_actions.py:37-68).  Returns true when anything changed. ----
// Everything it needs is passed by value: handing the kernel-parameter structs to a non-inlined function by reference
// would copy them to the caller's stack on the hot path.
struct DefenderCtx {
  uint8_t *left, *raw, *oo;                 // this env's reimage_left / owned_raw / owned_order rows
  uint32_t* changed;                        // this env's [words] "re-imaged in this step" set, or nullptr
  int32_t *n_raw, *n_owned, *n_reimaged;    // this env's scalars
  const int32_t* def_nodes;                 // this env's override rows (or nullptr)
  const float* def_uniforms;
  const uint8_t* reimageable;               // this scenario's nd_reimageable row
  int32_t* errflag;
  uint64_t seed, genv;
  double detect_prob;
  int words, ocap, scan_capacity, scan_frequency;
};
__device__ __noinline__ bool defender_step(DefenderCtx P, EnvBits M, int N, int stepcount, int total_steps, bool check_reown) {
  bool event = false;
  if (P.changed) for (int w = 0; w < P.words; ++w) P.changed[w] = 0u;
  uint8_t* left = P.left;
  uint8_t* raw = P.raw;
  int32_t& n_raw = *P.n_raw;
  // on_attacker_step_taken (:54-68): count down, then bring the node back (with the agent if it was persistent)
  for (int w = 0; w < P.words; ++w) {
    uint32_t img = M.word(M_IMAGING, w);
    while (img) {
      const int n = (w << 5) + __ffs(img) - 1;
      img &= img - 1;
      const int l = left[n];
      if (l > 0) left[n] = (uint8_t)(l - 1);
      else {
        M.clr(M_IMAGING, n);
        if (M.get(M_PERSISTENCE, n)) M.set(M_OWNED, n);
        event = check_reown = true;
      }
    }
  }
  // scan (:46-57): scan_capacity nodes drawn with replacement; a detection uniform is drawn only for eligible nodes
  if (stepcount % P.scan_frequency == 0) {
    int ui = 0, reimaged = 0;
    for (int j = 0; j < P.scan_capacity; ++j) {
      int n;
      if (P.def_nodes) n = P.def_nodes[j];
      else n = (int)(((uint64_t)philox_word(P.seed, P.genv, (uint32_t)total_steps, 3u, j) * (uint64_t)N) >> 32);
      if (n < 0 || n >= N) continue;
      if (M.get(M_STOPPED, n) || M.get(M_IMAGING, n) || !M.get(M_OWNED, n) || M.get(M_EVASION, n)) continue;
      float u;
      if (P.def_uniforms) u = P.def_uniforms[ui];
      else u = (float)(philox_word(P.seed, P.genv, (uint32_t)total_steps, 5u, ui) >> 8) * (1.0f / 16777216.0f);
      ++ui;
      if ((double)u <= P.detect_prob && P.reimageable[n]) {     // reimage_node (:37-52)
        left[n] = (uint8_t)REIMAGING_DURATION;
        M.clr(M_OWNED, n);
        M.set(M_IMAGING, n);
        M.set(M_OWN_STALE, n);                    // last_reimaging = now > last_owned_at
        if (P.changed) P.changed[n >> 5] |= 1u << (n & 31);
        ++reimaged;
        int k = 0;                                // owned_nodes.remove(node): first occurrence (cyberbattle_env.py:425)
        while (k < n_raw && raw[k] != n) ++k;
        if (k < n_raw) { for (; k + 1 < n_raw; ++k) raw[k] = raw[k + 1]; --n_raw; }
        // a persistent node that came back in this very step is not in the list yet (the re-own loop runs after the removals,
        // cyberbattle_env.py:425 vs :427-430): the reference raises ValueError here; counted, then treated as a no-op
        else atomicAdd(P.errflag + 1, 1);
        event = true;
      }
    }
    *P.n_reimaged += reimaged;
  }
  // :427-430 every Running node with the agent installed that is not in owned_nodes is appended, in node order
  if (check_reown) {
    uint32_t inl[MAX_NODES / 32] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < n_raw; ++k) inl[raw[k] >> 5] |= 1u << (raw[k] & 31);
    uint8_t* oo = P.oo;
    int32_t& n_owned = *P.n_owned;
    for (int w = 0; w < P.words; ++w) {
      uint32_t add = M.word(M_OWNED, w) & ~M.word(M_STOPPED, w) & ~M.word(M_IMAGING, w) & ~inl[w];
      while (add) {
        const int n = (w << 5) + __ffs(add) - 1;
        add &= add - 1;
        if (n_raw < P.ocap) raw[n_raw++] = (uint8_t)n; else atomicExch(P.errflag, 5);
        int k = 0;
        while (k < n_owned && oo[k] != n) ++k;
        if (k == n_owned) { oo[n_owned] = (uint8_t)n; ++n_owned; }
        event = true;
      }
    }
  }
  return event;
}

// ---- static defender step with ExternalRandomEvents (_env/static_defender.py:76-161): every node, in node order, draws one of
//      { start service, firewall remove, stop service, firewall add } and suffers it with probability `event_prob` unless it has
//      defense evasion.  Service events pick one of the node's services and (only while the node is Running) set its running
//      flag, but count as an event either way (static_defender_actions.py:150,161); firewall events pick one of the node's
//      ports, do nothing if an INCOMING rule (port, wanted permission) already exists — both sides test firewall.incoming,
//      :135-141 / :158-164 — else patch the rule of the drawn side (override_firewall_rule :96-128).  The device keeps one bit
//      per (node, service slot): running, incoming BLOCK, outgoing BLOCK (ev[node][0..2]).  Returns the number of events. ----
struct EventsCtx {
  uint16_t* ev;                 // this env's [ncap][4]
  const uint16_t* init;         // this scenario's nd_ev_init rows (service counts in [3])
  const float* draws;           // this env's [ncap][4] test override or nullptr
  uint64_t seed, genv;
  double prob;
};
__device__ __noinline__ int events_step(EventsCtx E, EnvBits M, int N, int total_steps) {
  int events = 0;
  for (int n = 0; n < N; ++n) {
    int f;
    float u_event, u_pick, u_side;
    if (E.draws) {
      const float4 d = reinterpret_cast<const float4*>(E.draws)[n];
      f = (int)d.x; u_event = d.y; u_pick = d.z; u_side = d.w;
    } else {      // Philox stream 16 + node: { function, event uniform, pick uniform, side uniform }
      const Philox4 r = philox4x32_10(E.seed, E.genv, (uint32_t)total_steps, 16u + (uint32_t)n);
      f = (int)(r.x >> 30);
      u_event = (float)(r.y >> 8) * (1.0f / 16777216.0f);
      u_pick = (float)(r.z >> 8) * (1.0f / 16777216.0f);
      u_side = (float)(r.w >> 8) * (1.0f / 16777216.0f);
    }
    if (M.get(M_EVASION, n) || !((double)u_event <= E.prob)) continue;
    const int ns = E.init[4 * n + 3];
    if (ns == 0) continue;                 // (the reference raises on a firewall event here; such scenarios are rejected up front)
    int slot = (int)(u_pick * (float)ns);
    if (slot > ns - 1) slot = ns - 1;
    uint16_t* w = E.ev + 4 * n;
    const uint16_t bit = (uint16_t)(1u << slot);
    if (f == 0 || f == 2) {
      if (!M.get(M_STOPPED, n) && !M.get(M_IMAGING, n)) w[0] = f == 0 ? (uint16_t)(w[0] | bit) : (uint16_t)(w[0] & ~bit);
      ++events;
    } else {
      const bool block = f == 3;
      if (((w[1] & bit) != 0) == block) continue;          // (port, permission) already among the incoming rules
      uint16_t& side = (double)u_side <= 0.5 ? w[1] : w[2];
      side = block ? (uint16_t)(side | bit) : (uint16_t)(side & ~bit);
      ++events;
    }
  }
  return events;
}

}  // namespace

// Everything a transition reads from the env's own records and the call's per-env inputs, requested in ONE burst before
// anything is consumed (the loads are independent; issuing them on demand made the kernel a chain of ~10 dependent
// round trips).  REG: one-word planes and no defender -> the mask record is staged in registers (EnvBitsReg).
template <bool REG>
struct TransitionIn {
  int4 h0, h1;            // hot sector of the scalar record: flags, stepcount, num_iterations, total steps | outcome, S_SCST, return
  int4 c0;                // first half of sector 1: n_disc, n_owned, disc_amount, n_owned_raw — needed by a third of the outcomes
                          // only, but fetched with the burst: on demand it is one more exposed DRAM round trip
  typename std::conditional<REG, EnvBitsReg, EnvBits>::type M;
  int4 sl;                // decoded action (source, target, vulnerability, outcome kind)
  double dist;
  float uniform;
  __device__ __forceinline__ void issue(const Params& P, const State& S, int b, const float* __restrict__ uniforms) {
    const int4* q = reinterpret_cast<const int4*>(S.scal + (size_t)b * 8);      // sector 0
    h0 = q[0]; h1 = q[1];
    c0 = *reinterpret_cast<const int4*>(S.scal + ((size_t)P.B + b) * 8);
    M.open(S.masks + (size_t)b * P.mpitch, P.words);
    uniform = uniforms ? uniforms[b] : 0.f;
  }
  __device__ __forceinline__ void store_hot(const State& S, int b) const {
    int4* q = reinterpret_cast<int4*>(S.scal + (size_t)b * 8);
    q[0] = h0; q[1] = h1;
  }
};

// DEF: a static defender is configured (compile-time, so the default kernels carry none of its code or registers)
// ENQ: the function itself appends the env to the decode cost bins and to the observe worklist (one atomic each: the
//      fused path, one lane per warp).  The thread-per-env kernel passes false and aggregates both per CTA — a million
//      same-address atomics serialise in L2 and were 90 % of that kernel's time at large batch.
// The new hot sector of the scalar record comes back in `in.h0 / in.h1` and the caller stores it as one full 32-byte
// sector (TransitionIn::store_hot); episode constants (reachable-node
// counts) come from the scenario tables, so sectors 2-3 of the record are never touched and sector 1 only when a list
// changes.  Table look-ups go level by level (scenario record -> instance index / node values -> instance record ->
// firewall word), every level's loads issued together and ahead of the validity chain that consumes them.
// Returns the env's observe work class (0 .. OBS_CLASSES-1, heaviest first) or -1.
// PERSIST: the caller keeps the env's records in registers across several steps (transition_ksteps_kernel): nothing of the
//      record is written here — mask words, list lengths and the hot sector all go back once, after the last step — and the
//      observe / decode work lists are not fed (no observe runs between those steps).
template <bool DEF, bool ENQ = true, bool REG = false, bool PERSIST = false>
static __device__ __forceinline__ int transition_env(const Tables& T, const Params& P, const State& S, int b, TransitionIn<REG>& in,
                                            const int4* __restrict__ sc_pack, bool have_uniform, bool write_sel, int sched_out,
                                            float* __restrict__ reward_out,
                                            uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out,
                                            uint8_t* __restrict__ outcome_out) {
  static_assert(!(DEF && REG), "the defender path works on the record in place");
  int32_t* cnt = S.scal + ((size_t)P.B + b) * 8;                   // sector 1 (list lengths / counters), on demand
  auto SC = [&](int plane) -> int32_t& { return cnt[plane - 8]; };
  // working copies of the list lengths (without a defender nothing else writes them during the step; the defender path
  // works on the words in place through DefenderCtx)
  int n_disc_w = in.c0.x, n_owned_w = in.c0.y, disc_amount_w = in.c0.z;
  const int4 h0 = in.h0, h1 = in.h1;
  auto& M = in.M;
  const int4 sl = in.sl;
  const double dist = in.dist;

  int flags = h0.x;
  if (ENQ) sched_enqueue(S, P, b, sched_out);   // cost-binned env list for the next decode (longest tables first)
  if (flags & (FL_DONE | FL_TRUNC | FL_NEEDS_RESET)) {
    // the reference raises RuntimeError here (cyberbattle_env.py:300-302); a finished env is left untouched
    if (reward_out) reward_out[b] = 0.f;
    if (done_out) done_out[b] = 1;
    if (trunc_out) trunc_out[b] = (flags & FL_TRUNC) ? 1 : 0;
    if (outcome_out) outcome_out[b] = OC_INVALID_SRC_NOT_OWNED;
    in.h0.x = flags & ~(FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP);
    return -1;
  }

  const int s = sl.x, t = sl.y, u = sl.z, kind = sl.w;
  if (write_sel) reinterpret_cast<int4*>(S.sel)[b] = sl;   // what the observe kernel and the info record read

  // ---- level 1: the scenario record ----
  const int sc = h1.y >> 8, starter = h1.y & 0xFF;        // S_SCST
  const int4 sp0 = sc_pack[2 * sc], sp1 = sc_pack[2 * sc + 1];
  const int N = sp0.x, node_off = sp0.y, U = sp0.z, port_off = sp0.w;
  const int64_t instof_off = ((int64_t)(uint32_t)sp1.x) | ((int64_t)sp1.y << 32);
  const int stepcount = h0.y + 1;                         // :303
  const int num_iter = h0.z;
  const int total_steps = h0.w;
  const bool local = (s == t);                            // :307
  const bool idx_ok = (s >= 0 && s < N && t >= 0 && t < N);
  const int tt = idx_ok ? t : 0, ss = idx_ok ? s : 0;     // safe indices for the look-ups issued ahead of the checks
  // ---- level 2: instance index, node values, reachable-node counts of (scenario, starter) — what reset_env keeps in
  //      S_OWNABLE / S_DISCOVERABLE / S_DISRUPTABLE / S_PROP_NODES, fetched from the tables (L2) so that the record's
  //      constant sector stays untouched ----
  const int bg = base_goal(P);
  const int32_t* reach_tab = bg == GOAL_CONTROL ? T.nd_ownable : (bg == GOAL_DISCOVERY ? T.nd_discoverable : T.nd_disruptable);
  const int reach = reach_tab[node_off + starter];
  int inst = -1;
  if (idx_ok && u >= 0 && u < U) inst = T.inst_of[instof_off + (int64_t)tt * U + u];
  const int t_value = T.nd_value[node_off + tt];
  const int t_laa = T.nd_level_at_access[node_off + tt];
  const float uf = have_uniform ? in.uniform
                                : philox_uniform(P.seed, (uint64_t)(P.global_env_offset + b), (uint32_t)total_steps, 0u);
  // ---- level 3: the instance record (flags | list lengths, outcome kinds, port, recon offset | success rate, cost) ----
  const int insti = inst >= 0 ? inst : 0;
  const uint4 vp0 = T.vi_pack[2 * insti], vp1 = T.vi_pack[2 * insti + 1];
  const uint32_t vf = vp0.x;
  uint2 recon_m = make_uint2(0u, 0u);      // node sets of the instance's two Reconnaissance lists (one-word planes only)
  if constexpr (REG) recon_m = T.recon_mask[insti];
  const double v_success = __hiloint2double((int)vp1.y, (int)vp1.x), v_cost = __hiloint2double((int)vp1.w, (int)vp1.z);
  // ---- outgoing-firewall word of (port, source): with one-word planes the instance record carries it ----
  uint32_t fw_out = 0u;
  bool listening = (vf & VI_LISTENING) != 0u, in_allowed = (vf & VI_IN_ALLOWED) != 0u;
  if (DEF && P.defender == 2) {
    // ExternalRandomEvents: services and firewall rules are per-env state (one bit per node and service slot)
    const uint16_t* ev = S.ev_cur + (size_t)b * P.ncap * 4;
    const int slot_t = (int)(vf >> 24);                                               // the target's service slot of the port
    const int slot_s = T.out_slot[(size_t)(port_off + (int)vp0.z) * T.max_nodes + ss];  // the source's
    listening = slot_t != 0xFF && ((ev[4 * tt] >> slot_t) & 1);
    in_allowed = slot_t == 0xFF || !((ev[4 * tt + 1] >> slot_t) & 1);
    fw_out = (slot_s != 0xFF && ((ev[4 * ss + 2] >> slot_s) & 1)) ? (1u << (ss & 31)) : 0u;
  } else {
    fw_out = P.words == 1 ? vp0.z : T.outblock[(size_t)(port_off + (int)vp0.z) * P.words + (ss >> 5)];
  }

  double reward = 0.0;
  int code = -1;

  // The validity chain and the per-outcome mutations below are written as straight-line selects wherever an arm is a bit test
  // plus a bit set: measured on the large-batch launch (ncu source view, round 2) the branchy version spent 50 % of the
  // kernel's warp instructions in these two sections at ~3 active lanes per warp (every lane of a warp wants another arm).
  // Only the arms with real work of their own (Reconnaissance list walk, privilege escalation, lateral move) stay branches.
  const bool own_s = M.get(M_OWNED, ss), disc_t = M.get(M_DISCOVERED, tt);
  const bool down_s = M.get(M_STOPPED, ss) || (DEF && M.get(M_IMAGING, ss));
  const bool down_t = M.get(M_STOPPED, tt) || (DEF && M.get(M_IMAGING, tt));
  const bool eva_s = M.get(M_EVASION, ss), eva_t = M.get(M_EVASION, tt);
  const bool root_t = M.get(M_PRIV_ROOT, tt), user_t = M.get(M_PRIV_USER, tt);
  const int level_t = root_t ? 3 : (user_t ? 1 : 0);
  {
    // ---- validity chain (attacker_actions.py:109-197 remote, :363-416 local): the FIRST failing test decides, so the tests
    //      are applied last to first and an earlier one overwrites a later one ----
    const int privreq = (vf >> VI_PRIVREQ_SHIFT) & 3;
    const uint32_t kinds = local ? (vp0.y & 0xFFFFu) : (vp0.y >> 16);
    int pidx = 0;
    if ((double)uf >= v_success) { code = OC_UNSUCCESSFUL; pidx = P_SUCCESS_FAILED; }                                    // :190 / :409
    if (!local && !eva_t && !in_allowed) { code = OC_FW_INCOMING; pidx = P_FW_REMOTE; }
    if (!local && !eva_s && ((fw_out >> (ss & 31)) & 1u)) { code = OC_FW_OUTGOING; pidx = P_FW_LOCAL; }
    if (!local && !listening) { code = OC_PORT_NOT_LISTENING; pidx = P_UNOPEN_PORT; }
    if (kind < 0 || kind >= N_KINDS || !((kinds >> (kind & 15)) & 1u)) { code = OC_OUTCOME_NOT_PRESENT; pidx = P_INVALID_ACTION; }
    if (privreq && level_t < privreq) { code = OC_NO_PRIVILEGE; pidx = P_NO_PRIV; }
    if (inst < 0) { code = OC_NO_VULNERABILITY; pidx = P_NO_VULN; }
    if (!local && down_t) { code = OC_TGT_NOT_RUNNING; pidx = P_INVALID_ACTION; }
    if (down_s) { code = OC_SRC_NOT_RUNNING; pidx = P_INVALID_ACTION; }
    if (!local && !disc_t) { code = OC_INVALID_TGT_NOT_DISCOVERED; pidx = P_INVALID_ACTION; }
    if (!idx_ok || !own_s) { code = OC_INVALID_SRC_NOT_OWNED; pidx = P_INVALID_ACTION; }
    if (code >= 0) reward = P.pen[pidx];
  }

  // ---- per-outcome mutation + reward (attacker_actions.py:199-351 / :418-547) ----
  if (code < 0) {
    double total = 0.0;
    bool ok = true;
    // the six outcomes that are "test a bit of the target, set a bit of the target" (t == tt here: the indices passed the chain)
    const bool kC = kind == K_COLLECTION, kP = kind == K_PERSISTENCE, kD = kind == K_DOS, kV = kind == K_DISCOVERY,
               kX = kind == K_EXFILTRATION, kE = kind == K_EVASION;
    if (kC | kP | kD | kV | kX | kE) {
      const bool okC = kC && M.get(M_HAS_DATA, t);
      const bool okP = kP && !M.get(M_PERSISTENCE, t);
      const bool okV = kV && !M.get(M_VISIBLE, t);
      const bool okX = kX && M.get(M_COLLECTED, t) && !M.get(M_EXFILTRATED, t);
      const bool okE = kE && !eva_t;
      // a stopped target never gets here (:127), a local DoS has no such test (:450): DoS always succeeds
      M.clr_if(M_HAS_DATA, t, okC); M.set_if(M_COLLECTED, t, okC);
      M.set_if(M_PERSISTENCE, t, okP);
      M.set_if(M_STOPPED, t, kD);
      M.set_if(M_VISIBLE, t, okV);
      M.set_if(M_EXFILTRATED, t, okX);
      M.set_if(M_EVASION, t, okE);
      ok = okC | okP | kD | okV | okX | okE;
      if (ok) {
        const int ridx = kC ? R_COLLECTED : (kP ? R_PERSISTENCE : (kD ? R_DOS : (kV ? R_VISIBILITY : (kX ? R_EXFILTRATED : R_EVASION))));
        total += kD ? P.rew[R_DOS] * (double)t_value : P.rew[ridx];
      } else {
        const int pidx = kC ? P_NO_DATA_COLLECT : (kP ? P_ALREADY_PERSISTENT : (kV ? P_ALREADY_VISIBLE : (kX ? P_NO_DATA_EXFIL : P_ALREADY_EVASION)));
        reward = P.pen[pidx];
        code = (kC | kX) ? OC_NO_NEEDED : OC_REPEATED;
      }
    } else switch (kind) {
      case K_RECON: {
        // the instance's two Reconnaissance lists sit in recon_pack, each starting on an 8-byte boundary: "any type"
        // first, then "REMOTE only".
        const int len_any = (vf >> 8) & 0xFF, len_remote = (vf >> 16) & 0xFF;
        const int off = (int)vp0.w + (local ? 0 : ((len_any + 7) & ~7)), len = local ? len_any : len_remote;
        int n_disc = DEF ? SC(S_N_DISC) : n_disc_w, fresh = 0;
        uint8_t* order = S.disc_order + (size_t)b * P.ncap;
        if constexpr (REG) {
          // one-word planes: the list's node set comes with the instance (recon_mask), so the nodes that are new are known
          // before the list is touched — most Reconnaissance outcomes discover nothing new and skip the walk; the others walk
          // it only until their last new node is appended (list order = the order of env.discovered_nodes, :291-296 +
          // cyberbattle_env.py:398-407)
          uint32_t todo = (local ? recon_m.x : recon_m.y) & ~M.word(M_DISCOVERED, 0);
          fresh = __popc(todo);
          M.or_word(M_DISCOVERED, todo);
          for (int i0 = 0; todo != 0u && i0 < len; i0 += 8) {
            const uint2 ch = *reinterpret_cast<const uint2*>(T.recon_pack + off + i0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int node = (int)(((j < 4 ? ch.x : ch.y) >> (8 * (j & 3))) & 0xFFu);
              if (i0 + j < len && ((todo >> (node & 31)) & 1u)) { todo &= ~(1u << (node & 31)); order[n_disc++] = (uint8_t)node; }
            }
          }
        } else {
          for (int i0 = 0; i0 < len; i0 += 8) {               // eight node ids per load; a node already discovered costs three instructions
            const uint2 ch = *reinterpret_cast<const uint2*>(T.recon_pack + off + i0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int node = (int)(((j < 4 ? ch.x : ch.y) >> (8 * (j & 3))) & 0xFFu);
              if (i0 + j < len && !M.get(M_DISCOVERED, node)) { M.set(M_DISCOVERED, node); order[n_disc++] = (uint8_t)node; ++fresh; }
            }
          }
        }
        if (DEF) { SC(S_N_DISC) = n_disc; SC(S_DISC_AMOUNT) += fresh; }
        else { n_disc_w = n_disc; disc_amount_w += fresh; }
        total += P.rew[R_NODE_DISCOVERED] * (double)fresh;
        break;
      }
      case K_PRIVESC: {
        const int lvl = (vf >> (local ? VI_LEVEL_ANY_SHIFT : VI_LEVEL_REMOTE_SHIFT)) & 3;
        const int cur = level_t;
        if (!local && cur == 0) { reward = P.pen[P_PRIVESC_NOT_OWNED]; code = OC_NO_PRIVILEGE; ok = false; }
        else if (cur >= lvl) { reward = P.pen[P_PRIVESC_ALREADY]; code = OC_REPEATED; ok = false; }
        else {                                              // __mark_node_as_owned(t, level) :70-89
          M.set(M_OWNED, t); M.set(M_DISCOVERED, t);
          if (DEF) { M.set(M_EVER_OWNED, t); M.clr(M_OWN_STALE, t); }
          if (lvl >= 1) M.set(M_PRIV_USER, t);
          if (lvl == 3) M.set(M_PRIV_ROOT, t);
          total += P.rew[R_PRIVESC];
        }
        break;
      }
      case K_LATERAL:
      case K_CREDACCESS:
        if (local) { reward = P.pen[P_OUTCOME_NOT_VALID]; code = OC_REMOTE_OUTCOME_LOCAL; ok = false; }  // :536-540
        else {
          // :327 marks before testing.  "Currently owned" = owned since the last re-imaging (:561-573); without a
          // defender that is the same as "ever owned" == agent installed
          bool ever = M.get(M_OWNED, t), was_owned = ever;
          if (DEF) {
            ever = M.get(M_EVER_OWNED, t);
            was_owned = ever && !M.get(M_OWN_STALE, t);
            M.set(M_EVER_OWNED, t); M.clr(M_OWN_STALE, t);
          }
          const int laa = t_laa;
          M.set(M_OWNED, t);
          if (laa >= 1) M.set(M_PRIV_USER, t);
          if (laa == 3) M.set(M_PRIV_ROOT, t);
          if (was_owned) { reward = P.pen[P_ALREADY_OWNED]; code = OC_REPEATED; ok = false; }
          else {
            if (!ever) total += P.rew[R_VALUE] * (double)t_value;   // :336-340 first ownership only
            uint8_t* oo = S.owned_order + (size_t)b * P.ncap;   // cyberbattle_env.py:408-410
            int n_owned = DEF ? SC(S_N_OWNED) : n_owned_w;
            if (DEF) {        // owned_nodes.append(t): the exact list may now hold t twice
              uint8_t* raw = S.owned_raw + (size_t)b * P.ocap;
              const int n_raw = SC(S_N_OWNED_RAW);
              if (n_raw < P.ocap) { raw[n_raw] = (uint8_t)t; SC(S_N_OWNED_RAW) = n_raw + 1; } else atomicExch(S.errflag, 5);
              int k = 0;
              while (k < n_owned && oo[k] != t) ++k;
              if (k == n_owned) { oo[n_owned] = (uint8_t)t; SC(S_N_OWNED) = n_owned + 1; }
            } else {
              oo[n_owned] = (uint8_t)t;
              n_owned_w = n_owned + 1;
            }
          }
        }
        break;
      default:  // Execution and anything else: :341-345
        reward = P.pen[P_OUTCOME_NOT_VALID]; code = OC_REMOTE_OUTCOME_LOCAL; ok = false;
        break;
    }
    if (ok) {
      total -= P.rew[R_COST] * v_cost;             // :348 / :544
      reward = total;
      code = kind;
      if (kC | kX | kV) {                                                                              // env:411-412
        if (DEF) SC(S_DISC_AMOUNT) += 1; else disc_amount_w += 1;
      }
    }
  }

  // ---- node-specific games: the reward is zeroed once the interest node is discovered and not targeted (:322-326) ----
  const int interest = is_node_goal(P) ? sp1.z : -1;
  if (interest >= 0 && t != interest && M.get(M_DISCOVERED, interest)) reward = 0.0;

  // ---- static defender (cyberbattle_env.py:331-332), then the cached-feature rule of the visible graph: the target's
  //      vector is rebuilt from the post-defender state for these obtained outcomes only (compressed:472-479) ----
  bool def_event = false;
  if constexpr (DEF) {
   if (P.defender == 2) {
    const int n_disc_before = in.c0.x;
    EventsCtx E;
    E.ev = S.ev_cur + (size_t)b * P.ncap * 4;
    E.init = T.nd_ev_init + (size_t)node_off * 4;
    E.draws = S.def_uniforms ? S.def_uniforms + (size_t)b * P.ncap * 4 : nullptr;
    E.seed = P.seed; E.genv = (uint64_t)(P.global_env_offset + b);
    E.prob = P.event_prob;
    const int ev_n = events_step(E, M, N, total_steps);
    SC(S_N_REIMAGED) += ev_n;                      // num_events of the episode (cyberbattle_env.py:419)
    def_event = ev_n > 0;
    // the visible graph caches node feature vectors: a node's services / firewall columns are those of the moment it joined
    // the graph (newly discovered nodes: after this step's events, compressed:467-469) or was last the target of a successful
    // outcome (compressed:472-479)
    uint16_t* evx = S.ev_x + (size_t)b * P.ncap * 4;
    const uint8_t* order = S.disc_order + (size_t)b * P.ncap;
    const int n_disc_now = SC(S_N_DISC);
    for (int i = n_disc_before; i < n_disc_now; ++i) {
      const int node = order[i];
      *reinterpret_cast<uint2*>(evx + 4 * node) = *reinterpret_cast<const uint2*>(E.ev + 4 * node);
    }
    if (code < 16 && code != K_RECON) *reinterpret_cast<uint2*>(evx + 4 * t) = *reinterpret_cast<const uint2*>(E.ev + 4 * t);
   } else {
    DefenderCtx D;
    D.left = S.reimage_left + (size_t)b * P.ncap;
    D.raw = S.owned_raw + (size_t)b * P.ocap;
    D.oo = S.owned_order + (size_t)b * P.ncap;
    D.changed = P.precise_positions ? S.changed + (size_t)b * P.words : nullptr;
    D.n_raw = &SC(S_N_OWNED_RAW); D.n_owned = &SC(S_N_OWNED); D.n_reimaged = &SC(S_N_REIMAGED);
    D.def_nodes = S.def_nodes ? S.def_nodes + (size_t)b * P.scan_capacity : nullptr;
    D.def_uniforms = S.def_uniforms ? S.def_uniforms + (size_t)b * P.scan_capacity : nullptr;
    D.reimageable = T.nd_reimageable + node_off;
    D.errflag = S.errflag;
    D.seed = P.seed; D.genv = (uint64_t)(P.global_env_offset + b);
    D.detect_prob = P.detect_prob;
    D.words = P.words; D.ocap = P.ocap; D.scan_capacity = P.scan_capacity; D.scan_frequency = P.scan_frequency;
    def_event = defender_step(D, M, N, stepcount, total_steps, code == K_PRIVESC);
    if (code < 16 && code != K_RECON) {
      if (M.get(M_IMAGING, t)) M.set(M_X_IMAGING, t); else M.clr(M_X_IMAGING, t);
     }
  }
  }

  // ---- goal / termination (cyberbattle_env.py:338-370, 438-514) ----
  int n_goal = 0, n_data = 0, n_pending = 0;
  bool any_running_owned = false;
  int roots_in_list = 0;
  bool interest_in_list = false;
  if (DEF) {   // the tests below run over env.owned_nodes, which is no longer the agent_installed set
    const uint8_t* raw = S.owned_raw + (size_t)b * P.ocap;
    const int n_raw = SC(S_N_OWNED_RAW);
    for (int k = 0; k < n_raw; ++k) {
      const int n = raw[k];
      if (!M.get(M_STOPPED, n) && !M.get(M_IMAGING, n)) any_running_owned = true;
      if (n != starter && M.get(M_PRIV_ROOT, n)) ++roots_in_list;        // duplicates count twice (:471-475)
      if (n == interest) interest_in_list = true;
    }
  }
  for (int w = 0; w < P.words; ++w) {     // one pass over the mask words: goal popcounts and the lost test
    const uint32_t own = M.word(M_OWNED, w), disc = M.word(M_DISCOVERED, w), stop = M.word(M_STOPPED, w);
    const uint32_t not_starter = ((starter >> 5) == w) ? ~(1u << (starter & 31)) : 0xFFFFFFFFu;
    if (P.goal == GOAL_CONTROL) n_goal += __popc(own & M.word(M_PRIV_ROOT, w) & not_starter);
    else if (P.goal == GOAL_DISRUPTION) n_goal += __popc(disc & stop & not_starter);
    else if (P.goal == GOAL_DISCOVERY) {
      n_goal += __popc(disc & not_starter);
      n_data += __popc(disc & M.word(M_HAS_DATA, w));
      n_pending += __popc(disc & M.word(M_COLLECTED, w) & ~M.word(M_EXFILTRATED, w));
    }
    if (!DEF) any_running_owned |= (own & ~stop) != 0u;
  }
  if (DEF && P.goal == GOAL_CONTROL) n_goal = roots_in_list;
  bool goal_ok;
  if (P.goal == GOAL_CONTROL || P.goal == GOAL_DISRUPTION) goal_ok = (n_goal == reach);
  else if (P.goal == GOAL_DISCOVERY) goal_ok = (n_goal == reach && n_data == 0 && n_pending == 0);
  // *_node goals (cyberbattle_env.py:467-514): a few bit tests on the interest node
  else if (P.goal == GOAL_CONTROL_NODE) goal_ok = (DEF ? interest_in_list : M.get(M_OWNED, interest)) && M.get(M_PRIV_ROOT, interest);
  else if (P.goal == GOAL_DISCOVERY_NODE)   // :493-508 (has_data is cleared by the collection, so "collected and exfiltrated" cannot hold with it)
    goal_ok = M.get(M_DISCOVERED, interest) && M.get(M_VISIBLE, interest) &&
              (!M.get(M_HAS_DATA, interest) || (M.get(M_COLLECTED, interest) && M.get(M_EXFILTRATED, interest)));
  else goal_ok = M.get(M_STOPPED, interest);
  // check_end_game (:438-454): killing the interest node loses control_node / discovery_node games
  const bool lost = ((P.goal == GOAL_CONTROL_NODE || P.goal == GOAL_DISCOVERY_NODE) && M.get(M_STOPPED, interest)) ||
                    !any_running_owned;

  int reason = 0;
  bool done = false, trunc = false;
  if (goal_ok) {
    if (P.goal == GOAL_DISRUPTION || P.stop_at_goal) done = true;
    reward = P.winning_reward; reason = 1;
  } else if (lost) {
    done = true; reward = P.losing_reward; reason = 2;
  } else if (P.prop_coeff != 0.0 && (double)num_iter >= (double)reach * P.prop_coeff) {
    trunc = true; reason = 3;
  } else if (num_iter >= P.episode_iterations) {
    trunc = true; reason = 3;
  }
  if (P.absolute_reward) reward = reward > 0.0 ? reward : 0.0;   // :379
  const bool add_edge = reward > 0.0;                            // compressed:483 (before the distance penalty)
  // compressed:401,462: the DESIRED outcome decides; with a defender or precise_graph_encoding every step re-encodes
  const bool reencode = P.always_encode || (kind == K_LATERAL || kind == K_DOS || kind == K_RECON);
  reward += P.pen[P_DISTANCE] * dist;                            // compressed:430

  // A re-encode of an unchanged graph reproduces the cached embeddings bit for bit, so it is skipped: `dirty`
  // records whether any node feature, edge or node set changed since the last encode (successful outcomes mutate
  // the target or the discovered set; reward > 0 adds / updates an edge).
  const bool dirty = (flags & FL_DIRTY) || code < 16 || add_edge || def_event;
  const int sticky = flags & FL_INTEREST_IN_GRAPH;          // survives until the episode's reset
  // (not with precise_action_space_positions: a re-encode of an unchanged graph still hands the current embeddings to
  // the pairs around THIS action's nodes, which an earlier refresh may have passed over)
  // (with sample_subset_samples the skipped build still advances the env's balance counter: it is counted in the flags word and
  // added by the next build that runs, or by the reset)
  const bool encode_now = reencode && (dirty || P.precise_positions);
  const int pending = ((h0.x >> FL_PENDING_SHIFT) & 0xFFFF) + ((P.subset_k && reencode && !encode_now) ? 1 : 0);
  flags = (done ? FL_DONE : 0) | (trunc ? FL_TRUNC : 0) | (reason << FL_REASON_SHIFT) | (add_edge ? FL_ADD_EDGE : 0) |
          (encode_now ? FL_REENCODE : 0) | ((done || trunc) ? FL_FINISHED_THIS_STEP : 0) | (dirty ? FL_DIRTY : 0) | sticky |
          (pending << FL_PENDING_SHIFT);
  if (!PERSIST) M.close();
  if (!DEF && !PERSIST) {   // changed list lengths go back word by word (fire and forget)
    if (n_disc_w != in.c0.x) SC(S_N_DISC) = n_disc_w;
    if (n_owned_w != in.c0.y) SC(S_N_OWNED) = n_owned_w;
    if (disc_amount_w != in.c0.z) SC(S_DISC_AMOUNT) = disc_amount_w;
  }
  if (PERSIST) { in.c0.x = n_disc_w; in.c0.y = n_owned_w; in.c0.z = disc_amount_w; }
  {   // the new hot sector (num_iterations += 1 is :394); the caller stores it
    const double ep = __hiloint2double(h1.w, h1.z) + reward;
    in.h0 = make_int4(flags, stepcount, num_iter + 1, total_steps + 1);
    in.h1 = make_int4(code, h1.y, __double2loint(ep), __double2hiint(ep));
  }
  int cls = -1;
  if (!PERSIST && (flags & (FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP))) {   // the observe kernel only visits these envs
    // OBS_CLASSES cost classes of 2.5 us, claimed heaviest first (tools/observe_trace.py, n = discovered nodes): a re-encode with
    // its table build costs about 9 + 0.85 n us, an episode end (statistics + reset + first observation) 13 us, an episode end
    // behind a re-encode (no table build) 20 + 0.65 n us, an edge update alone 3 us.  With six coarse classes an episode end
    // (13 us) was claimed before an 18-node re-encode (25 us), and the kernel ended a quarter later than its warps' mean load.
    const int n = DEF ? SC(S_N_DISC) : n_disc_w;
    const bool fin = flags & FL_FINISHED_THIS_STEP, enc = flags & FL_REENCODE;
    const int cost10 = fin ? (enc ? 200 + (13 * n) / 2 : 130) : (enc ? 90 + (17 * n) / 2 : 30);     // tenths of a microsecond
    const int bucket = cost10 / 25;
    cls = OBS_CLASSES - 1 - (bucket < OBS_CLASSES - 1 ? bucket : OBS_CLASSES - 1);
    if (ENQ) {
      const int slot = atomicAdd(&S.work_ctr[4 + cls], 1);
      if (slot < P.B) S.worklist[(size_t)cls * P.B + slot] = b; else atomicExch(S.errflag, 4);
    }
  }
  if (!PERSIST) S.reward64[b] = reward;
  if (reward_out) reward_out[b] = (float)reward;
  if (done_out) done_out[b] = (done || trunc) ? 1 : 0;
  if (trunc_out) trunc_out[b] = trunc ? 1 : 0;
  if (outcome_out) outcome_out[b] = (uint8_t)code;
  return cls;
}

}  // namespace cbs
