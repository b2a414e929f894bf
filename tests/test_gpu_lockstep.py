"""-m gpu: heterogeneous batch (different scenarios, sizes, action streams per env) in lockstep with one
oracle per env.  Covers per-env scenario offsets, the fused cbs_step call with auto-reset, on-device Philox
uniforms and random starters (both restated on the host from the same counters)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _philox_words(seed, env, step, stream):
    """Host restatement of csrc/philox.cuh (Philox4x32-10): the four output words."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    seed, env, step, stream = int(seed), int(env), int(step), int(stream)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c = [env & 0xFFFFFFFF, (env >> 32) & 0xFFFFFFFF, step & 0xFFFFFFFF, stream & 0xFFFFFFFF]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def _philox_uniform(seed, env, step, stream):
    return _philox_words(seed, env, step, stream)[0]


def philox_defender_draws(seed, env, step, n_nodes, capacity):
    """csrc/transition.cuh defender_step: scan draw j = word j&3 of stream 3 + (j>>2), scaled to [0, N); detection uniform i =
    word i&3 of stream 5 + (i>>2), 24-bit."""
    nodes = [(_philox_words(seed, env, step, 3 + (j >> 2))[j & 3] * n_nodes) >> 32 for j in range(capacity)]
    us = [np.float32((_philox_words(seed, env, step, 5 + (j >> 2))[j & 3] >> 8) * (1.0 / 16777216.0)) for j in range(capacity)]
    return nodes, us


def philox_events_draws(seed, env, step, n_nodes):
    """csrc/transition.cuh events_step: node n draws the four words of stream 16 + n = { function (top two bits), event uniform,
    pick uniform, side uniform (24-bit each) }"""
    rows = []
    for n in range(n_nodes):
        w = _philox_words(seed, env, step, 16 + n)
        rows.append([w[0] >> 30] + [float(np.float32((x >> 8) * (1.0 / 16777216.0))) for x in w[1:]])
    return (np.array(rows, np.float64),)


def philox_u01(seed, env, step, stream):
    return np.float32((_philox_uniform(seed, env, step, stream) >> 8) * (1.0 / 16777216.0))


def philox_pick(seed, env, step, stream, n):
    return (_philox_uniform(seed, env, step, stream) * n) >> 32


_DEFENDER = dict(static_defender_agent="reimage", detect_probability=0.7, scan_capacity=5, scan_frequency=2,
                 proportional_cutoff_coefficient=4)


@pytest.mark.parametrize("goal,sizes,pool_size,B,T,extra", [
    ("control", (8, 24), 120, 40, 70, {}),
    ("discovery", (8, 14), 120, 40, 70, {}),
    ("control", (40, 70), 120, 40, 70, {}),
    ("disruption", (6, 12), 60, 33, 60, {}),          # odd batch size
    ("control", (120, 128), 330, 4, 140, {}),         # maximum node count (4 mask words, 255 snapshot slots), Ug > 256: two GEMM N-tiles
    ("control", (3, 5), 30, 1, 80, {}),               # a single env, tiny scenarios
    ("control_node", (8, 16), 100, 24, 60, {}),       # node-specific goals: one interest node per scenario, 258-float observation
    ("discovery_node", (8, 16), 100, 24, 60, {}),
    ("control", (6, 14), 100, 32, 90, _DEFENDER),     # re-imaging defender on its Philox streams (3.. scan, 5.. detection)
    ("discovery", (8, 16), 100, 24, 60, dict(precise_graph_encoding=True)),
    ("control", (8, 20), 100, 24, 70, dict(precise_action_space_positions=True, proportional_cutoff_coefficient=3)),
    ("control", (8, 14), 100, 16, 60, dict(precise_action_space_positions=True, precise_graph_encoding=True)),
    # decode metrics other than cosine (compressed:571-576): k_decode_metric.cu + the transition as its own launch inside cbs_step
    ("control", (8, 24), 120, 40, 60, dict(distance_metric="l1")),
    ("discovery", (8, 20), 330, 33, 60, dict(distance_metric="l2")),        # Ug > 256: several tiles of the vulnerability part
    ("control", (20, 40), 120, 24, 50, dict(distance_metric="inf")),
    ("control", (6, 14), 100, 24, 60, dict(_DEFENDER, distance_metric="l2")),
    # sample_subset_samples (compressed:553-567): the reference's training default k = 100 on 32-node scenarios, a k small enough
    # to thin every class at every build, and the option together with the defender / precise_action_space_positions
    ("control", (28, 32), 120, 24, 70, dict(sample_subset_samples=100, proportional_cutoff_coefficient=2)),
    ("control", (8, 20), 100, 32, 70, dict(sample_subset_samples=6, proportional_cutoff_coefficient=3)),
    ("control", (6, 14), 100, 24, 80, dict(_DEFENDER, sample_subset_samples=10)),
    ("control", (8, 16), 100, 24, 70, dict(sample_subset_samples=8, precise_action_space_positions=True, proportional_cutoff_coefficient=3)),
    ("control", (8, 16), 100, 24, 60, dict(sample_subset_samples=8, distance_metric="l2", proportional_cutoff_coefficient=3)),
    # ExternalRandomEvents defender on its Philox streams (16 + node): services stopped / started, firewall rules added / removed
    ("control", (6, 14), 100, 24, 90, dict(static_defender_agent="events", random_event_probability=0.05, proportional_cutoff_coefficient=4)),
    # BASELINE configs[3] (bench workload c4): mixed 10-100-node scenarios in one padded batch, pool of 600 vulnerabilities
    ("control", (10, 100), 600, 18, 60, {}),
], ids=["control-8-24", "discovery-8-14", "control-40-70", "disruption-odd-batch", "control-128-nodes", "single-env-tiny",
        "control-node", "discovery-node", "defender-philox", "precise-encoding", "precise-positions", "precise-both",
        "metric-l1", "metric-l2", "metric-inf", "metric-l2-defender", "subset-k100-32-nodes", "subset-k6", "subset-defender",
        "subset-positions", "subset-l2", "events-philox", "c4-mixed-10-100"])
def test_lockstep_heterogeneous_batch(goal, sizes, pool_size, B, T, extra):
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from ccbs_b200.gae import GaeWeights
    from oracle import trace as tr
    from oracle.cbs_oracle import OracleEnv
    from tests.gpu_harness import masks_to_u64, near_tie_tolerance

    rng = np.random.default_rng(5)
    pool = cb.synthetic_vuln_pool(99, pool_size)
    S, seed, offset = (6 if (sizes[1] < 100 or sizes[0] < 50) else 2), 12345, 1000
    gkw = dict(vulns_per_service_range=(6, 14)) if pool_size > 256 else {}
    if sizes == (10, 100):      # the mixed batch: both ends of the range are present
        ns = [10, 100] + [int(x) for x in rng.integers(sizes[0], sizes[1] + 1, size=S - 2)]
    else:
        ns = [int(rng.integers(sizes[0], sizes[1] + 1)) for _ in range(S)]
    specs = [cb.synthetic_spec(200 + k, ns[k], pool=pool, **gkw) for k in range(S)]
    cfg = cb.EnvConfig(goal=goal, **extra)
    defender = cfg.static_defender_agent is not None
    w = GaeWeights.random(3)
    interest = None
    if goal.endswith("node"):      # an interest node that some starter can reach, per scenario
        interest, keep = [], []
        for sp in specs:
            probe = OracleEnv(sp, w, cfg, interest_node=0)
            cand = [i for i in range(sp.num_nodes) if any(i in probe._reach_sets[goal[:-5]][s] for s in range(sp.num_nodes) if s != i)]
            if cand:
                interest.append(int(cand[len(cand) // 2]))
                keep.append(sp)
        specs, S = keep, len(keep)
        assert S >= 2
    sc_of_env = rng.integers(0, S, size=B).astype(np.int32)
    env = BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, scenario_of_env=sc_of_env, seed=seed, global_env_offset=offset,
                                auto_reset=True, interest_nodes=interest)
    tables = env.tables
    g = cb.constants.GOALS[goal]
    G = env.obs_dim - 2
    oracles = [OracleEnv(specs[sc_of_env[b]], w, cfg, interest_node=None if interest is None else interest[sc_of_env[b]],
                         philox_seed=seed, env_index=offset + b) for b in range(B)]
    vidx = [tr.vuln_index(specs[sc_of_env[b]]) for b in range(B)]
    episodes = [0] * B
    total_steps = [0] * B

    def starter_for(b):
        sc = sc_of_env[b]
        f0, f1 = tables.sc_feasible_off[g][sc], tables.sc_feasible_off[g][sc + 1]
        return int(tables.feasible_starters[g][f0 + philox_pick(seed, offset + b, episodes[b], 1, int(f1 - f0))])

    obs = env.reset()
    env.sync()
    obs_h = obs.cpu().numpy()
    for b in range(B):
        o = oracles[b].reset(starter=starter_for(b))
        np.testing.assert_allclose(obs_h[b, :G], o["graph_embeddings"], rtol=1e-5, atol=2e-5)
    flips = 0
    actions = rng.uniform(-4, 4, size=(T, B, 905)).astype(np.float32)
    for t in range(T):
        obs, reward, done, info = env.step(torch.from_numpy(actions[t]).to(env.device), None)
        env.sync()
        obs_h, rew_h, done_h, info_h = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), info.cpu().numpy()
        m = env.masks()
        term = env.terminal_obs()
        for b in range(B):
            o = oracles[b]
            u = philox_u01(seed, offset + b, total_steps[b], 0)
            total_steps[b] += 1
            s, t_, vid, kind, d, _ = o.find_closest_action_embedding(actions[t, b])
            want = (s, t_, vidx[b][vid], kind)
            got = tuple(int(x) for x in info_h[b, :4])
            forced = None
            if got != want:
                # accepted only as a genuine near-tie of the ORACLE's float64 distances (two table rows whose
                # node embeddings agree to the last float32 bit or so); the oracle then follows the CUDA pick
                dd = o.all_distances(actions[t, b])
                cand = [i for i, k in enumerate(o.action_keys) if (k[0], k[1], vidx[b][k[2]], k[3]) == got]
                assert cand, f"step {t} env {b}: decode {got} is not in the oracle's table (oracle {want})"
                i = min(cand, key=lambda j: dd[j])
                assert dd[i] - d < near_tie_tolerance(cfg.distance_metric), f"step {t} env {b}: decode {got} (gap {dd[i] - d:.3e}) vs oracle {want}"
                flips += 1
                forced = (got[0], got[1], o.action_keys[i][2], got[3], dd[i])
            if cfg.static_defender_agent == "events":
                dd_ = philox_events_draws(seed, offset + b, total_steps[b] - 1, o.N)
            else:
                dd_ = philox_defender_draws(seed, offset + b, total_steps[b] - 1, o.N, int(cfg.scan_capacity)) if defender else None
            ob, r, dn, inf = o.step(actions[t, b], u, forced=forced, defender_draws=dd_)
            assert int(info_h[b, 4]) == o.outcome and int(info_h[b, 5]) == o.end_episode_reason
            assert bool(done_h[b]) == bool(dn)
            np.testing.assert_allclose(rew_h[b], r, rtol=1e-5, atol=1e-4)
            if dn:
                np.testing.assert_allclose(term[b, :G], ob["graph_embeddings"], rtol=1e-5, atol=2e-5)
                assert tuple(term[b, G:]) == tuple(float(x) for x in ob["discrete_features"])
                episodes[b] += 1
                ob = o.reset(starter=starter_for(b))
            np.testing.assert_allclose(obs_h[b, :G], ob["graph_embeddings"], rtol=1e-5, atol=2e-5)
            assert tuple(obs_h[b, G:]) == tuple(float(x) for x in ob["discrete_features"])
            assert np.array_equal(masks_to_u64(m, b), tr.masks_to_array(o.masks())), f"step {t} env {b}: masks differ"
    sc = env.scalars()
    # l1 sums the float32 rounding differences of equal embeddings instead of cancelling them: wider near-tie window, more of them
    assert flips <= max(1, B * T // (50 if cfg.distance_metric == "l1" else 200)), f"{flips} near-tie flips in {B * T} env-steps"
    assert np.array_equal(sc[L.S_EPISODES], np.array(episodes))
    assert sum(episodes) >= (B if sizes[1] < 100 else 1), "test too short to exercise auto-reset"
    if cfg.sample_subset_samples:
        dropped = sum(o.rows_dropped for o in oracles)
        print("rows dropped by the sub-sampling:", dropped, "balance calls:", sum(o.balance_calls for o in oracles))
        assert dropped > 0, "the sub-sampling never thinned a class: k too large for this case"
    print(f"near-tie flips: {flips} in {B * T} env-steps")
    if pool_size > 256:
        assert tables.vemb32.shape[0] > 256
    acc = env.stat_accum()
    assert acc["episodes"] == sum(episodes)
    if cfg.distance_metric == "cosine":
        assert env.margin_edge_count() == 0, "a decode came near the edge of the float32 re-score margin"
    if cfg.static_defender_agent == "reimage":
        assert acc["stat9"] > 0 and acc["lost"] > 0, "the defender never re-imaged a node / never evicted the attacker"
    if cfg.static_defender_agent == "events":
        assert acc["stat10"] > 0 and acc["stat9"] == 0, "no external event happened"
    env.close()
