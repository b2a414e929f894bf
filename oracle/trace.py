"""TEST INFRASTRUCTURE — trace protocol shared by the golden generator (reference side) and the
oracle / GPU parity tests.

A trace is: a scenario, GAE weights (seeded), an env config, a float32 action sequence, a float32
uniform sequence (success-rate draws) and one starter node per episode.  The env is stepped with
auto-reset (reset right after a step that returned done-or-truncated, as DummyVecEnv does) and after
every step the full integer state, the reward, the flags, the decoded action and the observation are
recorded.
"""
from __future__ import annotations

import numpy as np

import ccbs_b200.constants as C


def make_inputs(seed: int, steps: int):
    """Actions U(-4,4)^905 float32 (the action_space bounds, compressed:112-114) and uniforms in [0,1)."""
    rng = np.random.default_rng(seed)
    actions = rng.uniform(-4.0, 4.0, size=(steps, C.ACTION_DIM)).astype(np.float32)
    uniforms = rng.random(steps, dtype=np.float32)
    return actions, uniforms


def make_defender_draws(seed: int, steps: int, num_nodes: int, capacity: int):
    """Per step: `capacity` node indices (random.choices, static_defender.py:48) and `capacity` detection uniforms
    (numpy.random.random, :53; consumed in call order)."""
    rng = np.random.default_rng(seed)
    return (rng.integers(num_nodes, size=(steps, capacity)).astype(np.int32),
            rng.random((steps, capacity), dtype=np.float32))


def make_events_draws(seed: int, steps: int, num_nodes: int, p_event: float, boost: float = 8.0):
    """ExternalRandomEvents draws, one row (f, u_event, u_pick, u_side) per step and node (see OracleEnv.events_defender_step).
    u_event is skewed towards the event probability (a fraction ``boost * p_event`` of the draws fall below it) so that a short
    trace sees many events; the defender itself is unchanged."""
    rng = np.random.default_rng(seed)
    ev = np.zeros((steps, num_nodes, 4), np.float64)
    ev[..., 0] = rng.integers(0, 4, size=(steps, num_nodes))
    hit = rng.random((steps, num_nodes)) < min(1.0, boost * p_event)
    ev[..., 1] = np.where(hit, rng.random((steps, num_nodes)) * p_event, p_event + (1 - p_event) * (0.001 + 0.999 * rng.random((steps, num_nodes))))
    ev[..., 2] = rng.random((steps, num_nodes))
    ev[..., 3] = rng.random((steps, num_nodes))
    return (ev,)


def make_starters(seed: int, feasible, count: int):
    rng = np.random.default_rng(seed)
    feasible = np.asarray(feasible)
    return feasible[rng.integers(len(feasible), size=count)].astype(np.int32)


def vuln_index(spec):
    """vulnerability_ID -> scenario-local unique index (first-appearance order, same as the compiler)."""
    idx = {}
    for nd in spec.nodes:
        for v in nd.vulns:
            idx.setdefault(v.vid, len(idx))
    return idx


def masks_to_array(masks):
    out = np.zeros((C.N_MASKS, 2), dtype=np.uint64)
    for i, m in enumerate(masks):
        out[i, 0] = m & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = (m >> 64) & 0xFFFFFFFFFFFFFFFF
    return out


class OracleAdapter:
    def __init__(self, env, spec):
        self.env, self.vidx, self.N = env, vuln_index(spec), spec.num_nodes

    def reset(self, starter):
        o = self.env.reset(starter=int(starter))
        return np.concatenate([o["graph_embeddings"].astype(np.float32), o["discrete_features"].astype(np.float32)])

    def step(self, action, u, forced=None, defender_draws=None):
        o, r, d, info = self.env.step(action, u, forced=forced, defender_draws=defender_draws)
        obs = np.concatenate([o["graph_embeddings"].astype(np.float32), o["discrete_features"].astype(np.float32)])
        e = self.env
        sel = (info["source_node"], info["target_node"], self.vidx[info["vulnerability"]], info["outcome_kind"])
        return obs, float(r), bool(e.done), bool(e.truncated), sel, int(e.outcome), int(e.end_episode_reason), \
            float(info["min_distance_action"])

    def masks(self):
        return self.env.masks()

    def table_keys(self):
        return [(k[0], k[1], k[3]) for k in self.env.action_keys]

    def table_row(self, i):
        return np.asarray(self.env.action_rows[i], dtype=np.float64)

    def owned_and_root(self):
        e = self.env
        return set(e.owned_nodes), {n for n in e.owned_nodes if e.nodes[n].privilege_level == C.PRIV_ROOT}

    def lists(self):
        return list(self.env.discovered_nodes), list(self.env.owned_nodes)

    def counters(self):
        e = self.env
        return (e.stepcount, e.num_iterations, e.discovered_amount, e.ownable_count, e.discoverable_count,
                e.disruptable_count, e.discoverable_amount)

    def statistics(self):
        return self.env.get_statistics()


class ReferenceAdapter:
    """Same surface over oracle.ref_bridge.ReferenceRunner (build container only)."""

    def __init__(self, runner, spec):
        self.r, self.vidx, self.N = runner, vuln_index(spec), spec.num_nodes

    @staticmethod
    def _obs(o):
        return np.concatenate([np.asarray(o["graph_embeddings"], dtype=np.float32),
                               np.asarray(o["discrete_features"], dtype=np.float32)])

    def reset(self, starter):
        return self._obs(self.r.reset(starter))

    def step(self, action, u, forced=None, defender_draws=None):
        from ccbs_b200.scenario import _KIND_BY_CLASSNAME
        o, reward, done_or_trunc, truncated, info = self.r.step(action, u, defender_draws)
        env = self.r.env
        kind = _KIND_BY_CLASSNAME[type(info["outcome_class"]).__name__]
        sel = (self.r.index[info["source_node"]], self.r.index[info["target_node"]], self.vidx[info["vulnerability"]], kind)
        assert bool(done_or_trunc) == bool(env.done or env.truncated)
        return self._obs(o), float(reward), bool(env.done), bool(env.truncated), sel, self.r.obtained_code(), \
            int(info["end_episode_reason"]), float(info["min_distance_action"])

    def masks(self):
        return self.r.masks()

    def table_keys(self):
        from ccbs_b200.scenario import _KIND_BY_CLASSNAME
        ix = self.r.index
        return [(ix[k[0]], ix[k[1]], _KIND_BY_CLASSNAME[type(k[3]).__name__]) for k in self.r.env.action_embeddings]

    def table_row(self, i):
        return np.asarray(list(self.r.env.action_embeddings.values())[i], dtype=np.float64)

    def owned_and_root(self):
        e, ix = self.r.env, self.r.index
        return ({ix[n] for n in e.owned_nodes},
                {ix[n] for n in e.owned_nodes if int(e.get_node(n).privilege_level) == C.PRIV_ROOT})

    def lists(self):
        e = self.r.env
        return [self.r.index[n] for n in e.discovered_nodes], [self.r.index[n] for n in e.owned_nodes]

    def counters(self):
        e = self.r.env
        return (e.stepcount, e.num_iterations, e.discovered_amount, e.ownable_count, e.discoverable_count,
                e.disruptable_count, e.discoverable_amount)

    def statistics(self):
        return self.r.wrapper.get_statistics()


def policy_pick(keys, owned, root, rng, p_greedy=0.85, p_persist=0.0):
    """Scripted attacker used for the 'policy' golden cases: prefer lateral moves onto nodes not yet owned, then
    privilege escalation on owned non-root nodes, then reconnaissance; otherwise (and with prob. 1 - p_greedy) a
    uniformly random table row.  ``keys`` = [(source, target, kind)] in table order.  Returns a row index."""
    if p_persist and rng.random() < p_persist:      # defender cases: make owned nodes persistent so re-imaging hands them back
        pers = [i for i, (s, t, k) in enumerate(keys) if k == C.K_PERSISTENCE and t in owned]
        if pers:
            return int(pers[rng.integers(len(pers))])
    if rng.random() < p_greedy:
        lateral = [i for i, (s, t, k) in enumerate(keys) if k == C.K_LATERAL and t not in owned]
        privesc = [i for i, (s, t, k) in enumerate(keys) if k == C.K_PRIVESC and t in owned and t not in root]
        recon = [i for i, (s, t, k) in enumerate(keys) if k == C.K_RECON]
        for cls, p in ((lateral, 0.6), (privesc, 0.8), (recon, 0.5)):
            if cls and rng.random() < p:
                return int(cls[rng.integers(len(cls))])
    return int(rng.integers(len(keys)))


def record(adapter, actions, uniforms, starters, policy_seed=None, policy_rows=None, defender_draws=None, p_persist=0.0):
    """Step ``adapter`` through the whole action sequence with auto-reset; return a dict of arrays.

    Policy mode (``policy_seed`` set): ``actions`` holds small float32 noise only; the action of step t is a row of the
    env's current action table plus that noise.  When ``policy_rows`` is None the row is chosen by :func:`policy_pick`
    (generation, reference side) and recorded; otherwise the recorded rows are replayed (oracle / CUDA side)."""
    T, N = len(actions), adapter.N
    OW = N if defender_draws is None else 2 * N      # under a defender env.owned_nodes can hold duplicates (env:425-430)
    prng = np.random.default_rng(policy_seed) if policy_seed is not None else None
    chosen_rows = np.zeros(T, np.int32)
    rec = dict(sel=np.zeros((T, 4), np.int32), code=np.zeros(T, np.int32), reward=np.zeros(T, np.float64),
               done=np.zeros(T, np.uint8), truncated=np.zeros(T, np.uint8), reason=np.zeros(T, np.uint8),
               dist=np.zeros(T, np.float64), masks=np.zeros((T, C.N_MASKS, 2), np.uint64),
               disc_order=np.full((T, N), -1, np.int16), owned_order=np.full((T, OW), -1, np.int16),
               counters=np.zeros((T, 7), np.int32), obs=None, episode=np.zeros(T, np.int32))
    reset_obs, reset_masks, stats = [], [], []
    ep = 0
    reset_obs.append(adapter.reset(starters[ep]))
    rec["obs"] = np.zeros((T, len(reset_obs[0])), np.float32)      # 194, or 258 for *_node goals
    reset_masks.append(masks_to_array(adapter.masks()))
    for t in range(T):
        action = actions[t]
        if policy_seed is not None:
            if policy_rows is None:
                owned, root = adapter.owned_and_root()
                chosen_rows[t] = policy_pick(adapter.table_keys(), owned, root, prng, p_persist=p_persist)
            else:
                chosen_rows[t] = policy_rows[t]
            action = (adapter.table_row(int(chosen_rows[t])) + actions[t].astype(np.float64)).astype(np.float32)
        dd = None if defender_draws is None else tuple(x[t] for x in defender_draws)
        obs, r, done, trunc, sel, code, reason, dist = adapter.step(action, uniforms[t], defender_draws=dd)
        rec["sel"][t], rec["code"][t], rec["reward"][t] = sel, code, r
        rec["done"][t], rec["truncated"][t], rec["reason"][t], rec["dist"][t] = done, trunc, reason, dist
        rec["masks"][t] = masks_to_array(adapter.masks())
        d, o = adapter.lists()
        rec["disc_order"][t, :len(d)] = d
        rec["owned_order"][t, :len(o)] = o
        rec["counters"][t] = adapter.counters()
        rec["obs"][t] = obs
        rec["episode"][t] = ep
        if done or trunc:
            stats.append([float(x) for x in adapter.statistics()])
            ep += 1
            if ep >= len(starters):
                raise RuntimeError("not enough starters for the number of episodes")
            reset_obs.append(adapter.reset(starters[ep]))
            reset_masks.append(masks_to_array(adapter.masks()))
    rec["reset_obs"] = np.array(reset_obs, np.float32)
    rec["reset_masks"] = np.array(reset_masks, np.uint64)
    rec["stats"] = np.array(stats, np.float64).reshape(-1, 14)
    rec["num_episodes"] = np.array(ep + 1, np.int32)
    if policy_seed is not None:
        rec["policy_rows"] = chosen_rows
    return rec


def window(rec, actions, uniforms, starters, first_episode, max_steps):
    """Cut a recorded trace at an episode boundary: the steps of episodes ``first_episode``.. (at most ``max_steps`` of
    them) with the matching inputs, as a self-contained (record, actions, uniforms, starters).  Episodes are independent
    given their starter, so a replay of the window must reproduce it exactly (random-action traces without a defender)."""
    t0 = int(np.argmax(rec["episode"] >= first_episode))
    if rec["episode"][t0] != first_episode:
        raise ValueError("first_episode is beyond the trace")
    t1 = min(len(rec["episode"]), t0 + int(max_steps))
    out = {}
    for k, v in rec.items():
        if k in ("reset_obs", "reset_masks", "stats", "num_episodes"):
            continue
        out[k] = np.array(v[t0:t1], copy=True)
    out["episode"] = out["episode"] - first_episode
    k = int(np.count_nonzero(out["done"] | out["truncated"]))
    out["reset_obs"] = rec["reset_obs"][first_episode:first_episode + k + 1]
    out["reset_masks"] = rec["reset_masks"][first_episode:first_episode + k + 1]
    out["stats"] = rec["stats"][first_episode:first_episode + k]
    out["num_episodes"] = np.array(k + 1, np.int32)
    return out, actions[t0:t1], uniforms[t0:t1], starters[first_episode:]


INT_KEYS = ("sel", "code", "done", "truncated", "reason", "masks", "disc_order", "owned_order", "counters",
            "episode", "reset_masks", "num_episodes")   # "policy_rows" is an input, not compared


def compare(a, b, rtol=1e-5, atol=1e-5, label=""):
    """Integer state bit-exact; reward / distance / obs / stats within tolerance.  Returns a report."""
    report = {}
    for k in INT_KEYS:
        if k in ("masks", "reset_masks") and a[k].shape[-2] != b[k].shape[-2]:
            # fixtures recorded before the defender planes existed: the extra planes must be empty, the rest equal
            n = min(a[k].shape[-2], b[k].shape[-2])
            for x in (a[k], b[k]):
                if np.any(x[..., n:, :]):
                    raise AssertionError(f"{label}: '{k}' has defender-plane bits set but the other side does not record them")
            a, b = dict(a), dict(b)
            a[k], b[k] = a[k][..., :n, :], b[k][..., :n, :]
        if not np.array_equal(a[k], b[k]):
            bad = np.argwhere(np.asarray(a[k]) != np.asarray(b[k]))
            raise AssertionError(f"{label}: integer field '{k}' differs first at {bad[0].tolist()} "
                                 f"({np.asarray(a[k])[tuple(bad[0])]} vs {np.asarray(b[k])[tuple(bad[0])]})")
    for k in ("reward", "dist", "obs", "reset_obs", "stats"):
        x, y = np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)
        if x.shape != y.shape:
            raise AssertionError(f"{label}: '{k}' shape {x.shape} vs {y.shape}")
        err = np.abs(x - y)
        tol = atol + rtol * np.abs(y)
        report[k] = float(err.max()) if err.size else 0.0
        if np.any(err > tol):
            i = np.unravel_index(np.argmax(err - tol), err.shape)
            raise AssertionError(f"{label}: '{k}' differs at {i}: {x[i]} vs {y[i]} (|d|={err[i]:.3e})")
    return report
