"""Per-step CSV trace of selected envs, in the column layout of RandomSwitchEnv.add_to_csv
(_env/cyberbattle_env_switch.py:223-279: the variant without embeddings).

The reference writes one row per step from Python object state; here the integer state of the traced envs is read back from
the device before and after each step (a debugging aid: it synchronises, keep the traced subset small).  :func:`write_replay_csv`
writes the same rows for T steps from the device log of ``BatchedCyberBattleEnv.replay`` (``cbs_replay``): one synchronisation
for the whole trace.  Node detail
strings follow get_str_info (switch.py:307-334) character for character (tests/test_switch_rule.py compares them with the
reference's own on a scripted-attacker trace); "Iteration" is the 0-based step of the episode (the reference writes the row before
it increments steps_in_current_episode, switch.py:124-134).  Deliberate differences: the two outcome columns hold class names
where the reference prints object reprs with memory addresses (switch.py:259-264,276), and every step of the traced envs is
written, whereas the reference skips episode 0 and writes every save_to_csv_interval-th episode (switch.py:124)."""
from __future__ import annotations

import csv
from typing import Sequence

import numpy as np

from . import constants as C
from . import lib as L

HEADER = ["Environment", "Episode", "Iteration", "Discovered Nodes", "Owned Nodes", "Alive nodes", "Source node",
          "Target node", "Vulnerability ID", "Outcome Mapped", "Reward", "Outcome", "Done", "Source Node Details",
          "Target Node Details", "Previous Source Node Details", "Previous Target Node Details", "Edges"]
# PrivilegeLevel is an IntEnum: str() is the bare integer under the reference's Python (3.12, environment.yml:22)
_PRIV = {0: "0", 1: "1", 3: "3"}


class TraceCsvWriter:
    def __init__(self, env, path: str, env_ids: Sequence[int] = (0,)):
        self.env, self.ids = env, [int(i) for i in env_ids]
        self.file = open(path, "w", newline="")
        self.writer = csv.writer(self.file)
        self.writer.writerow(HEADER)
        self.edges = {b: [] for b in self.ids}
        self._before = None

    # ---- state snapshots -------------------------------------------------------------------------
    def _snapshot(self):
        env = self.env
        env.sync()
        return dict(masks=env.masks(), scal=env.scalars(), disc=env.disc_order(), owned=env.owned_order())

    def _node_str(self, snap, b, node):
        t = self.env.tables
        sc = int(snap["scal"][L.S_SCENARIO, b])
        nd = t.specs[sc].nodes[node]
        bit = lambda plane: bool((snap["masks"][plane, node >> 5, b] >> (node & 31)) & 1)   # noqa: E731
        priv = 3 if bit(C.M_PRIV_ROOT) else (1 if bit(C.M_PRIV_USER) else 0)
        s = (f"status : {'MachineStatus.Stopped' if bit(C.M_STOPPED) else 'MachineStatus.Running'} / tag : {nd.tag} / value : "
             f"{nd.value} / privilege level : {_PRIV[priv]} / has data : {bit(C.M_HAS_DATA)} / data collected : "
             f"{bit(C.M_COLLECTED)} / data exfiltrated : {bit(C.M_EXFILTRATED)} / visible : {bit(C.M_VISIBLE)} / persistence : "
             f"{bit(C.M_PERSISTENCE)} / defense evasion : {bit(C.M_EVASION)} / level at access : {_PRIV[nd.level_at_access]} / services : ")
        for svc in nd.services:
            fin = int(any(p == svc.port and perm == 0 for p, perm in nd.fw_in))
            fout = int(any(p == svc.port and perm == 0 for p, perm in nd.fw_out))
            s += f"{svc.port} {svc.running} {fin} {fout} "
        s += " / vulnerabilities : "
        for v in nd.vulns:
            s += f"{v.vid}  " + "".join(f"{'remote' if r.vtype else 'local'}--{C.KIND_LABELS[r.kind]}  " for r in v.results)
        return s

    def _lists(self, snap, b):
        t = self.env.tables
        sc = int(snap["scal"][L.S_SCENARIO, b])
        ids = t.node_ids[sc]
        nd, no = int(snap["scal"][L.S_N_DISC, b]), int(snap["scal"][L.S_N_OWNED, b])
        disc = [ids[j] for j in snap["disc"][b, :nd]]
        owned = [ids[j] for j in snap["owned"][b, :no]]
        alive = [ids[j] for j in snap["disc"][b, :nd] if not (snap["masks"][C.M_STOPPED, j >> 5, b] >> (j & 31)) & 1]
        return sc, ids, disc, owned, alive

    # ---- per-step protocol -----------------------------------------------------------------------
    def before_step(self):
        self._before = self._snapshot()

    def after_step(self, reward, done, info):
        """reward[B], done[B], info[B,8] as returned by BatchedCyberBattleEnv.step (tensors or arrays)."""
        to_np = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)   # noqa: E731
        reward, done, info = to_np(reward), to_np(done), to_np(info)
        after, before = self._snapshot(), self._before
        t = self.env.tables
        for b in self.ids:
            sc, ids, disc, owned, alive = self._lists(before, b)
            s, tg, u, kind, code, _, step_count, _ = (int(x) for x in info[b])
            finished = bool(done[b])
            # after an in-place reset the post-step node state is gone: the details columns then show the fresh episode
            src_now = self._node_str(after, b, s) if int(after["scal"][L.S_SCENARIO, b]) == sc else ""
            tgt_now = self._node_str(after, b, tg) if int(after["scal"][L.S_SCENARIO, b]) == sc else ""
            if float(reward[b]) > 0 and not finished:    # an edge is added when the step's reward is positive (compressed:483)
                self.edges[b].append(f"{ids[s]}:{ids[tg]}:{t.vuln_ids[sc][u]}")
            self.writer.writerow([sc, int(before["scal"][L.S_EPISODES, b]), step_count - 1, disc, owned, alive, ids[s], ids[tg],
                                  t.vuln_ids[sc][u], C.KIND_NAMES[kind], float(reward[b]),
                                  C.KIND_NAMES[code] if code < 16 else C.OC_NAMES.get(code), finished, src_now, tgt_now,
                                  self._node_str(before, b, s), self._node_str(before, b, tg), ",".join(self.edges[b])])
            if finished:
                self.edges[b] = []
        self.file.flush()

    def close(self):
        self.file.close()


def write_replay_csv(env, log: dict, path: str, env_ids: Sequence[int] = (0,), first_env: int = 0):
    """The rows :class:`TraceCsvWriter` writes step by step, for all T steps of a ``BatchedCyberBattleEnv.replay`` log (the device
    wrote every step's records into it; nothing is read back per step).  ``env_ids`` index the envs of the batch, ``first_env`` is
    the first env the log covers.  The replay must have started right after a reset (episode starts are recovered from the
    logged post-reset masks: a fresh episode's lists are [starter]).  Row for row identical to the step-by-step writer."""
    t = env.tables
    W = t.words
    helper = TraceCsvWriter.__new__(TraceCsvWriter)
    helper.env = env
    T = log["sel"].shape[0]

    def snap(masks_e, disc, owned, nd, no, sc, episode):      # the snapshot layout TraceCsvWriter's helpers read, one env at index 0
        scal = np.zeros((L.NUM_SCALARS, 1), np.int64)
        scal[L.S_SCENARIO, 0], scal[L.S_N_DISC, 0], scal[L.S_N_OWNED, 0], scal[L.S_EPISODES, 0] = sc, nd, no, episode
        return dict(masks=np.ascontiguousarray(masks_e.reshape(C.N_MASKS, W, 1)), scal=scal, disc=disc[None, :], owned=owned[None, :])

    def fresh(masks_e, sc, episode):                          # state right after a reset: the starter alone
        own = masks_e[C.M_OWNED]
        starter = next(32 * w + int(own[w]).bit_length() - 1 for w in range(W) if own[w])
        lst = np.zeros(env.ncap, np.uint8)
        lst[0] = starter
        return snap(masks_e, lst, lst, 1, 1, sc, episode)

    with open(path, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(HEADER)
        for b in env_ids:
            e = int(b) - first_env
            edges = []
            before = None
            for k in range(T):
                sc, ep = int(log["scenario"][k, e]), int(log["episode"][k, e])
                if before is None:
                    # first step of the log: the pre-step state is a fresh episode's (has_data / visible come from the scenario,
                    # the starter is the head of the step's own owned list)
                    m0 = np.zeros((C.N_MASKS, W), np.uint32)
                    m0[C.M_HAS_DATA], m0[C.M_VISIBLE] = t.sc_init_has_data[sc], t.sc_init_visible[sc]
                    starter = int(log["owned_order"][k, e, 0])
                    bit = np.uint32(1 << (starter % 32))
                    laa = int(t.nd_level_at_access[t.sc_node_off[sc] + starter])
                    for plane in (C.M_OWNED, C.M_DISCOVERED) + ((C.M_PRIV_USER,) if laa >= 1 else ()) + ((C.M_PRIV_ROOT,) if laa == 3 else ()):
                        m0[plane, starter // 32] |= bit
                    before = fresh(m0, sc, ep)
                s, tg, u, kind = (int(x) for x in log["sel"][k, e])
                code, finished = int(log["code"][k, e]), bool(log["done"][k, e] or log["truncated"][k, e])
                after = snap(log["masks"][k, e], log["disc_order"][k, e], log["owned_order"][k, e], int(log["n_disc"][k, e]),
                             int(log["n_owned"][k, e]), sc, ep)
                _, ids, disc, owned, alive = helper._lists(before, 0)
                nxt = fresh(log["reset_masks"][k, e], int(log["scenario"][k + 1, e]) if k + 1 < T else sc, ep + 1) if finished else after
                same = int(nxt["scal"][L.S_SCENARIO, 0]) == sc
                reward = float(np.float32(log["reward"][k, e]))
                if reward > 0 and not finished:
                    edges.append(f"{ids[s]}:{ids[tg]}:{t.vuln_ids[sc][u]}")
                wr.writerow([sc, ep, int(log["step_count"][k, e]) - 1, disc, owned, alive, ids[s], ids[tg], t.vuln_ids[sc][u],
                             C.KIND_NAMES[kind], reward, C.KIND_NAMES[code] if code < 16 else C.OC_NAMES.get(code), finished,
                             helper._node_str(nxt, 0, s) if same else "", helper._node_str(nxt, 0, tg) if same else "",
                             helper._node_str(before, 0, s), helper._node_str(before, 0, tg), ",".join(edges)])
                if finished:
                    edges = []
                before = nxt
