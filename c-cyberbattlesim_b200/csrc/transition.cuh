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

namespace cbs {

namespace {

struct EnvBits {
  uint32_t* masks;
  int words, B, b;
  __device__ uint32_t word(int plane, int w) const { return masks[((size_t)plane * words + w) * B + b]; }
  __device__ bool get(int plane, int node) const { return (word(plane, node >> 5) >> (node & 31)) & 1u; }
  __device__ void set(int plane, int node) { masks[((size_t)plane * words + (node >> 5)) * B + b] |= (1u << (node & 31)); }
  __device__ void clr(int plane, int node) { masks[((size_t)plane * words + (node >> 5)) * B + b] &= ~(1u << (node & 31)); }
};

}  // namespace

static __device__ __forceinline__ void transition_env(const Tables& T, const Params& P, const State& S, int b, int4 sl, double dist,
                                            const float* __restrict__ uniforms, int sched_out, float* __restrict__ reward_out,
                                            uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out,
                                            uint8_t* __restrict__ outcome_out) {
  const int B = P.B;
  int32_t* scal = S.scal;
  auto SC = [&](int plane) -> int32_t& { return scal[(size_t)plane * B + b]; };

  int flags = SC(S_FLAGS);
  sched_enqueue(S, P, b, sched_out);   // cost-binned env list for the next decode (longest tables first)
  if (flags & (FL_DONE | FL_TRUNC | FL_NEEDS_RESET)) {
    // the reference raises RuntimeError here (cyberbattle_env.py:300-302); a finished env is left untouched
    if (reward_out) reward_out[b] = 0.f;
    if (done_out) done_out[b] = 1;
    if (trunc_out) trunc_out[b] = (flags & FL_TRUNC) ? 1 : 0;
    if (outcome_out) outcome_out[b] = OC_INVALID_SRC_NOT_OWNED;
    SC(S_FLAGS) = flags & ~(FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP);
    return;
  }

    const int s = sl.x, t = sl.y, u = sl.z, kind = sl.w;
  reinterpret_cast<int4*>(S.sel)[b] = sl;

  const int sc = SC(S_SCENARIO);
  const int N = T.sc_num_nodes[sc];
  const int node_off = T.sc_node_off[sc];
  EnvBits M{S.masks, P.words, B, b};

  SC(S_STEPCOUNT) += 1;                                   // :303
  const int total_steps = SC(S_TOTAL_STEPS);
  SC(S_TOTAL_STEPS) = total_steps + 1;

  double reward = 0.0;
  int code = -1;
  const bool local = (s == t);                            // :307
  const bool idx_ok = (s >= 0 && s < N && t >= 0 && t < N);
  int inst = -1;
  uint32_t vf = 0;

  // ---- validity chain (attacker_actions.py:109-197 remote, :363-416 local) ----
  if (!idx_ok || !M.get(M_OWNED, s)) {
    reward = P.pen[P_INVALID_ACTION]; code = OC_INVALID_SRC_NOT_OWNED;
  } else if (!local && !M.get(M_DISCOVERED, t)) {
    reward = P.pen[P_INVALID_ACTION]; code = OC_INVALID_TGT_NOT_DISCOVERED;
  } else if (M.get(M_STOPPED, s)) {
    reward = P.pen[P_INVALID_ACTION]; code = OC_SRC_NOT_RUNNING;
  } else if (!local && M.get(M_STOPPED, t)) {
    reward = P.pen[P_INVALID_ACTION]; code = OC_TGT_NOT_RUNNING;
  } else {
    const int U = T.sc_num_uvuln[sc];
    if (u >= 0 && u < U) inst = T.inst_of[T.sc_instof_off[sc] + (int64_t)t * (U > 0 ? U : 1) + u];
    if (inst < 0) {
      reward = P.pen[P_NO_VULN]; code = OC_NO_VULNERABILITY;
    } else {
      vf = T.vi_flags[inst];
      const int privreq = (vf >> VI_PRIVREQ_SHIFT) & 3;
      const int level = M.get(M_PRIV_ROOT, t) ? 3 : (M.get(M_PRIV_USER, t) ? 1 : 0);
      const uint32_t kinds = local ? T.vi_kinds_any[inst] : T.vi_kinds_remote[inst];
      if (privreq && level < privreq) {
        reward = P.pen[P_NO_PRIV]; code = OC_NO_PRIVILEGE;
      } else if (kind < 0 || kind >= N_KINDS || !((kinds >> kind) & 1u)) {
        reward = P.pen[P_INVALID_ACTION]; code = OC_OUTCOME_NOT_PRESENT;
      } else if (!local && !(vf & VI_LISTENING)) {
        reward = P.pen[P_UNOPEN_PORT]; code = OC_PORT_NOT_LISTENING;
      } else if (!local && !M.get(M_EVASION, s) &&
                 ((T.outblock[(size_t)(T.sc_port_off[sc] + T.vi_port[inst]) * P.words + (s >> 5)] >> (s & 31)) & 1u)) {
        reward = P.pen[P_FW_LOCAL]; code = OC_FW_OUTGOING;
      } else if (!local && !M.get(M_EVASION, t) && !(vf & VI_IN_ALLOWED)) {
        reward = P.pen[P_FW_REMOTE]; code = OC_FW_INCOMING;
      } else {
        const float uf = uniforms ? uniforms[b]
                                  : philox_uniform(P.seed, (uint64_t)(P.global_env_offset + b), (uint32_t)total_steps, 0u);
        if ((double)uf >= T.vi_success[inst]) {            // :190 / :409
          reward = P.pen[P_SUCCESS_FAILED]; code = OC_UNSUCCESSFUL;
        }
      }
    }
  }

  // ---- per-outcome mutation + reward (attacker_actions.py:199-351 / :418-547) ----
  if (code < 0) {
    double total = 0.0;
    bool ok = true;
    switch (kind) {
      case K_COLLECTION:
        if (M.get(M_HAS_DATA, t)) { M.clr(M_HAS_DATA, t); M.set(M_COLLECTED, t); total += P.rew[R_COLLECTED]; }
        else { reward = P.pen[P_NO_DATA_COLLECT]; code = OC_NO_NEEDED; ok = false; }
        break;
      case K_PERSISTENCE:
        if (M.get(M_PERSISTENCE, t)) { reward = P.pen[P_ALREADY_PERSISTENT]; code = OC_REPEATED; ok = false; }
        else { M.set(M_PERSISTENCE, t); total += P.rew[R_PERSISTENCE]; }
        break;
      case K_DOS:  // a stopped target never gets here (:127), a local DoS has no such test (:450)
        M.set(M_STOPPED, t); total += P.rew[R_DOS] * (double)T.nd_value[node_off + t];
        break;
      case K_DISCOVERY:
        if (M.get(M_VISIBLE, t)) { reward = P.pen[P_ALREADY_VISIBLE]; code = OC_REPEATED; ok = false; }
        else { M.set(M_VISIBLE, t); total += P.rew[R_VISIBILITY]; }
        break;
      case K_EXFILTRATION:
        if (M.get(M_COLLECTED, t) && !M.get(M_EXFILTRATED, t)) { M.set(M_EXFILTRATED, t); total += P.rew[R_EXFILTRATED]; }
        else { reward = P.pen[P_NO_DATA_EXFIL]; code = OC_NO_NEEDED; ok = false; }
        break;
      case K_EVASION:
        if (M.get(M_EVASION, t)) { reward = P.pen[P_ALREADY_EVASION]; code = OC_REPEATED; ok = false; }
        else { M.set(M_EVASION, t); total += P.rew[R_EVASION]; }
        break;
      case K_RECON: {
        const int32_t* rl = local ? T.vi_recon_any : T.vi_recon_remote;
        const int off = rl[2 * inst], len = rl[2 * inst + 1];
        int n_disc = SC(S_N_DISC), fresh = 0;
        uint8_t* order = S.disc_order + (size_t)b * P.ncap;
        for (int i = 0; i < len; ++i) {                     // :291-296 + cyberbattle_env.py:398-407
          const int node = T.recon_nodes[off + i];
          if (!M.get(M_DISCOVERED, node)) { M.set(M_DISCOVERED, node); order[n_disc++] = (uint8_t)node; ++fresh; }
        }
        SC(S_N_DISC) = n_disc;
        SC(S_DISC_AMOUNT) += fresh;
        total += P.rew[R_NODE_DISCOVERED] * (double)fresh;
        break;
      }
      case K_PRIVESC: {
        const int lvl = (vf >> (local ? VI_LEVEL_ANY_SHIFT : VI_LEVEL_REMOTE_SHIFT)) & 3;
        const int cur = M.get(M_PRIV_ROOT, t) ? 3 : (M.get(M_PRIV_USER, t) ? 1 : 0);
        if (!local && cur == 0) { reward = P.pen[P_PRIVESC_NOT_OWNED]; code = OC_NO_PRIVILEGE; ok = false; }
        else if (cur >= lvl) { reward = P.pen[P_PRIVESC_ALREADY]; code = OC_REPEATED; ok = false; }
        else {                                              // __mark_node_as_owned(t, level) :70-89
          M.set(M_OWNED, t); M.set(M_DISCOVERED, t);
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
          const bool was_owned = M.get(M_OWNED, t);         // :327 marks before testing
          const int laa = T.nd_level_at_access[node_off + t];
          M.set(M_OWNED, t);
          if (laa >= 1) M.set(M_PRIV_USER, t);
          if (laa == 3) M.set(M_PRIV_ROOT, t);
          if (was_owned) { reward = P.pen[P_ALREADY_OWNED]; code = OC_REPEATED; ok = false; }
          else {
            total += P.rew[R_VALUE] * (double)T.nd_value[node_off + t];
            uint8_t* oo = S.owned_order + (size_t)b * P.ncap;   // cyberbattle_env.py:408-410
            int n_owned = SC(S_N_OWNED);
            oo[n_owned] = (uint8_t)t;
            SC(S_N_OWNED) = n_owned + 1;
          }
        }
        break;
      default:  // Execution and anything else: :341-345
        reward = P.pen[P_OUTCOME_NOT_VALID]; code = OC_REMOTE_OUTCOME_LOCAL; ok = false;
        break;
    }
    if (ok) {
      total -= P.rew[R_COST] * T.vi_cost[inst];             // :348 / :544
      reward = total;
      code = kind;
      if (kind == K_COLLECTION || kind == K_EXFILTRATION || kind == K_DISCOVERY) SC(S_DISC_AMOUNT) += 1;  // env:411-412
    }
  }

  // ---- node-specific games: the reward is zeroed once the interest node is discovered and not targeted (:322-326) ----
  const int interest = is_node_goal(P) ? T.sc_interest[sc] : -1;
  if (interest >= 0 && t != interest && M.get(M_DISCOVERED, interest)) reward = 0.0;

  // ---- goal / termination (cyberbattle_env.py:338-370, 438-514) ----
  const int starter = SC(S_STARTER);
  int n_goal = 0, n_data = 0, n_pending = 0;
  bool any_running_owned = false;
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
    any_running_owned |= (own & ~stop) != 0u;
  }
  bool goal_ok;
  if (P.goal == GOAL_CONTROL) goal_ok = (n_goal == SC(S_OWNABLE));
  else if (P.goal == GOAL_DISRUPTION) goal_ok = (n_goal == SC(S_DISRUPTABLE));
  else if (P.goal == GOAL_DISCOVERY) goal_ok = (n_goal == SC(S_DISCOVERABLE) && n_data == 0 && n_pending == 0);
  else goal_ok = goal_reached(S, P, b, interest);     // *_node goals: a few bit tests on the interest node
  // check_end_game (:438-454): killing the interest node loses control_node / discovery_node games
  const bool lost = ((P.goal == GOAL_CONTROL_NODE || P.goal == GOAL_DISCOVERY_NODE) && M.get(M_STOPPED, interest)) ||
                    !any_running_owned;

  const int num_iter = SC(S_NUM_ITER);
  int reason = 0;
  bool done = false, trunc = false;
  if (goal_ok) {
    if (P.goal == GOAL_DISRUPTION || P.stop_at_goal) done = true;
    reward = P.winning_reward; reason = 1;
  } else if (lost) {
    done = true; reward = P.losing_reward; reason = 2;
  } else if (P.prop_coeff != 0.0 && (double)num_iter >= (double)SC(S_PROP_NODES) * P.prop_coeff) {
    trunc = true; reason = 3;
  } else if (num_iter >= P.episode_iterations) {
    trunc = true; reason = 3;
  }
  if (P.absolute_reward) reward = reward > 0.0 ? reward : 0.0;   // :379
  const bool add_edge = reward > 0.0;                            // compressed:483 (before the distance penalty)
  const bool reencode = (kind == K_LATERAL || kind == K_DOS || kind == K_RECON);   // compressed:401,462 (desired outcome)
  reward += P.pen[P_DISTANCE] * dist;                            // compressed:430
  SC(S_NUM_ITER) = num_iter + 1;                                 // :394
  SC(S_OUTCOME) = code;

  // A re-encode of an unchanged graph reproduces the cached embeddings bit for bit, so it is skipped: `dirty`
  // records whether any node feature, edge or node set changed since the last encode (successful outcomes mutate
  // the target or the discovered set; reward > 0 adds / updates an edge).
  const bool dirty = (flags & FL_DIRTY) || code < 16 || add_edge;
  const int sticky = flags & FL_INTEREST_IN_GRAPH;          // survives until the episode's reset
  const bool encode_now = reencode && dirty;
  flags = (done ? FL_DONE : 0) | (trunc ? FL_TRUNC : 0) | (reason << FL_REASON_SHIFT) | (add_edge ? FL_ADD_EDGE : 0) |
          (encode_now ? FL_REENCODE : 0) | ((done || trunc) ? FL_FINISHED_THIS_STEP : 0) | (dirty ? FL_DIRTY : 0) | sticky;
  SC(S_FLAGS) = flags;
  if (flags & (FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP)) {   // the observe kernel only visits these envs
    // three cost classes, claimed heaviest first: episode end (statistics + reset + encode + table), re-encode, edge only
    const int cls = (flags & FL_FINISHED_THIS_STEP) ? 0 : ((flags & FL_REENCODE) ? 1 : 2);
    const int slot = atomicAdd(&S.work_ctr[4 + cls], 1);
    if (slot < P.B) S.worklist[(size_t)cls * P.B + slot] = b; else atomicExch(S.errflag, 4);
  }
  S.reward64[b] = reward;
  S.ep_return[b] += reward;
  if (reward_out) reward_out[b] = (float)reward;
  if (done_out) done_out[b] = (done || trunc) ? 1 : 0;
  if (trunc_out) trunc_out[b] = trunc ? 1 : 0;
  if (outcome_out) outcome_out[b] = (uint8_t)code;
}

}  // namespace cbs
