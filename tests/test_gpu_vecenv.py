"""-m gpu: the SB3-VecEnv-shaped adapter and the single-env gymnasium-shaped wrapper."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vec_env_protocol():
    import ccbs_b200 as cb
    from ccbs_b200.vec_env import CyberBattleVecEnv
    specs = [cb.synthetic_spec(500 + k, 10) for k in range(3)]
    benv = cb.BatchedCyberBattleEnv(specs, cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=24, seed=3)
    venv = CyberBattleVecEnv(benv)
    obs = venv.reset()
    assert obs["graph_embeddings"].shape == (24, 192) and obs["discrete_features"].shape == (24, 2)
    assert obs["graph_embeddings"].dtype == np.float64 and np.all(obs["discrete_features"] == 1)
    rng = np.random.default_rng(0)
    episodes, ret = 0, np.zeros(24)
    for t in range(40):
        a = rng.uniform(-4, 4, size=(24, 905)).astype(np.float32)
        obs, rew, done, infos = venv.step(a)
        assert rew.shape == (24,) and done.dtype == bool and len(infos) == 24
        ret += rew
        for b in range(24):
            i = infos[b]
            assert i["source_node"] in benv.tables.node_ids[benv.scenario_of_env[b]]
            assert i["vulnerability_type"] in ("local", "remote") and i["outcome"] is not None
            if done[b]:
                episodes += 1
                assert i["terminal_observation"]["graph_embeddings"].shape == (192,)
                assert i["TimeLimit.truncated"] is False and i["end_episode_reason"] in (1, 2, 3)
                assert len(i["episode_stats"]) == 14 and i["episode_stats"][4] == 10
                assert abs(i["episode"]["r"] - ret[b]) < 1e-2 * max(1.0, abs(ret[b])) and i["episode"]["l"] >= 1
                ret[b] = 0
                assert obs["discrete_features"][b, 0] == 1          # already reset
    assert episodes > 24
    stats = venv.env_method("get_statistics")
    assert len(stats) == 24 and len(stats[0]) == 14
    assert venv.env_is_wrapped(object) == [False] * 24
    venv.close()


def test_single_env_wrapper_raises_after_done():
    import ccbs_b200 as cb
    from ccbs_b200.vec_env import RandomSwitchEnvB200
    env = RandomSwitchEnvB200([cb.synthetic_spec(7, 9)], cb.GaeWeights.random(0), cb.EnvConfig())
    obs, info = env.reset()
    assert obs["graph_embeddings"].shape == (192,) and info == {}
    rng = np.random.default_rng(1)
    for t in range(60):
        obs, r, done, trunc, info = env.step(rng.uniform(-4, 4, size=905).astype(np.float32))
        assert 0.0 <= info["min_distance_action"] <= 2.0
        if done:
            break
    assert done and len(env.get_statistics()) == 14
    with pytest.raises(RuntimeError):
        env.step(np.zeros(905, np.float32))
    obs, _ = env.reset()
    assert obs["discrete_features"][0] == 1
    env.close()


def test_csv_trace_writer(tmp_path):
    import csv
    import torch
    import ccbs_b200 as cb
    from ccbs_b200.trace_csv import TraceCsvWriter, HEADER
    env = cb.BatchedCyberBattleEnv([cb.synthetic_spec(11, 9)], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=6, seed=2)
    env.reset()
    path = tmp_path / "logs.csv"
    w = TraceCsvWriter(env, str(path), env_ids=(0, 4))
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    for t in range(25):
        w.before_step()
        obs, rew, done, info = env.step(torch.rand(6, 905, device="cuda", generator=g) * 8 - 4, None)
        w.after_step(rew, done, info)
    w.close()
    rows = list(csv.reader(open(path)))
    assert rows[0] == HEADER and len(rows) == 1 + 2 * 25
    assert all(len(r) == len(HEADER) for r in rows)
    assert {r[11] for r in rows[1:]} & {"Reconnaissance", "RepeatedResult", "InvalidAction", "Discovery", "NoNeededAction"}
    assert "status : MachineStatus.Running" in rows[1][15]
    env.close()
    # the same rows from cbs_replay's device log: one synchronisation for the whole trace instead of two state read-backs per step
    from ccbs_b200.trace_csv import write_replay_csv
    env = cb.BatchedCyberBattleEnv([cb.synthetic_spec(11, 9)], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=6, seed=2)
    env.reset()
    g.manual_seed(0)
    actions = torch.stack([torch.rand(6, 905, device="cuda", generator=g) * 8 - 4 for _ in range(25)]).contiguous()
    log = env.replay(actions, None)
    path2 = tmp_path / "logs_replay.csv"
    write_replay_csv(env, log, str(path2), env_ids=(0, 4))
    rows2 = list(csv.reader(open(path2)))
    by_env = lambda rr: sorted(rr[1:], key=lambda r: 0)      # noqa: E731  (the step-by-step writer interleaves the envs, this one does not)
    assert rows2[0] == HEADER and len(rows2) == len(rows)
    step_rows = {0: rows[1::2], 4: rows[2::2]}
    replay_rows = {0: rows2[1:26], 4: rows2[26:51]}
    for b in (0, 4):
        for k, (r1, r2) in enumerate(zip(step_rows[b], replay_rows[b])):
            assert r1 == r2, (b, k, [(h, x, y) for h, x, y in zip(HEADER, r1, r2) if x != y])
    env.close()


def test_device_vec_normalize_runs():
    import torch
    import ccbs_b200 as cb
    from ccbs_b200.normalize import DeviceVecNormalize
    env = cb.BatchedCyberBattleEnv([cb.synthetic_spec(12, 10)], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=128, seed=2)
    vn = DeviceVecNormalize(env)
    obs = vn.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    for t in range(20):
        obs, rew, done, _ = vn.step(torch.rand(128, 905, device="cuda", generator=g) * 8 - 4)
    assert obs.shape == (128, 194) and torch.isfinite(obs).all() and obs.abs().max() <= 10.0
    assert torch.isfinite(rew).all() and rew.abs().max() <= 10.0 and float(vn.ret_rms.count) > 2000
    env.close()


def test_sharded_host_env_matches_single_handle():
    """K handles stepped as a host pipeline give exactly what one handle holding the whole batch gives."""
    import torch
    import ccbs_b200 as cb
    specs = [cb.synthetic_spec(800 + k, 11) for k in range(3)]
    w, cfg = cb.GaeWeights.random(2), cb.EnvConfig()
    B, T = 50, 45
    rng = np.random.default_rng(5)
    actions = rng.uniform(-4, 4, size=(T, B, 905)).astype(np.float32)

    def run(env):
        env.reset()
        env.sync()
        pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt).pin_memory()  # noqa: E731
        a, obs, rew, done, info = pin(B, 905), pin(B, 194), pin(B), pin(B, dt=torch.uint8), pin(B, 8, dt=torch.int32)
        out = []
        for t in range(T):
            a.numpy()[...] = actions[t]
            env.step_host(a.numpy(), None, obs.numpy(), rew.numpy(), done.numpy(), info.numpy())
            out.append((obs.numpy().copy(), rew.numpy().copy(), done.numpy().copy(), info.numpy().copy()))
        return out, env.terminal_obs(), env.last_stats(), env.stat_accum()

    one = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, seed=41)
    ref, ref_term, ref_stats, ref_acc = run(one)
    one.close()
    for shards in (3, 4):
        env = cb.ShardedHostEnv(specs, w, cfg, num_envs=B, shards=shards, seed=41)
        assert [hi - lo for lo, hi in env.bounds] == [len(x) for x in np.array_split(np.arange(B), shards)]
        got, term, stats, acc = run(env)
        env.close()
        for t in range(T):
            for x, y in zip(got[t], ref[t]):
                assert np.array_equal(x, y), f"shards {shards} step {t}"
        assert np.array_equal(term, ref_term) and np.array_equal(stats, ref_stats)
        assert acc["episodes"] == ref_acc["episodes"] > B and abs(acc["return_sum"] - ref_acc["return_sum"]) < 1e-6 * abs(ref_acc["return_sum"])


def test_vec_env_over_sharded_host_env():
    """The SB3 adapter gives the same rollout over the pipelined host env as over one handle."""
    import ccbs_b200 as cb
    from ccbs_b200.vec_env import CyberBattleVecEnv
    specs = [cb.synthetic_spec(520 + k, 9) for k in range(2)]
    w, cfg = cb.GaeWeights.random(0), cb.EnvConfig()
    rng = np.random.default_rng(8)
    actions = rng.uniform(-4, 4, size=(30, 20, 905)).astype(np.float32)

    def rollout(venv):
        out = [venv.reset()]
        for a in actions:
            obs, rew, done, infos = venv.step(a)
            out.append((obs, rew, done, [(i["source_node"], i["target_node"], i["vulnerability"], i["outcome"], i["end_episode_reason"],
                                          i.get("episode_stats")) for i in infos]))
        venv.close()
        return out

    a = rollout(CyberBattleVecEnv(cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=20, seed=5)))
    b = rollout(CyberBattleVecEnv(cb.ShardedHostEnv(specs, w, cfg, num_envs=20, shards=3, seed=5)))
    for k in ("graph_embeddings", "discrete_features"):
        assert np.array_equal(a[0][k], b[0][k])
    for (oa, ra, da, ia), (ob, rb, db, ib) in zip(a[1:], b[1:]):
        assert np.array_equal(oa["graph_embeddings"], ob["graph_embeddings"]) and np.array_equal(ra, rb) and np.array_equal(da, db)
        assert ia == ib


def test_state_view_and_episode_stats_calls():
    """cbs_get_state hands out the same device pointers as cbs_state_ptr; cbs_episode_stats copies the 14-tuples."""
    import ctypes as ct
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    env = cb.BatchedCyberBattleEnv([cb.synthetic_spec(31, 10)], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=12, seed=4)
    env.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    for _ in range(25):
        env.step(torch.rand(12, 905, device="cuda", generator=g) * 8 - 4, None)
    env.sync()
    view = L.CbsStateView()
    assert env.lib.cbs_get_state(env._h, ct.byref(view)) == 0
    assert (view.num_envs, view.words, view.obs_dim, view.num_masks) == (12, 1, 194, cb.constants.N_MASKS)
    for name, field in (("masks", L.F_MASKS), ("scalars", L.F_SCALARS), ("obs", L.F_OBS), ("sel", L.F_SEL), ("dist", L.F_DIST),
                        ("last_stats", L.F_LAST_STATS), ("stat_accum", L.F_STAT_ACCUM), ("pair_slot", L.F_PAIR_SLOT)):
        assert getattr(view, name) == env.lib.cbs_state_ptr(env._h, field), name
    stats = np.empty((12, 14), np.float64)
    assert env.lib.cbs_episode_stats(env._h, stats.ctypes.data_as(ct.c_void_p)) == 0
    assert np.array_equal(stats, env.last_stats()) and stats[:, 4].max() == 10
    env.close()


def test_info_dicts_follow_on_device_scenario_switches():
    """with switch_interval the device draws a new scenario at resets; the info dicts must name nodes / vulnerabilities
    of the scenario in force during the step (carried in the info row), not of the env's initial assignment."""
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    from ccbs_b200.vec_env import CyberBattleVecEnv
    specs = [cb.synthetic_spec(900 + k, 7 + 2 * k) for k in range(4)]
    for k, sp in enumerate(specs):                       # node ids that tell the scenarios apart
        for nd in sp.nodes:
            nd.node_id = f"s{k}:{nd.node_id}"
    B = 16
    env = cb.BatchedCyberBattleEnv(specs, cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=9, switch_interval=1,
                                   scenario_of_env=np.zeros(B, np.int32))
    venv = CyberBattleVecEnv(env, lazy_infos=False)
    venv.reset()
    rng = np.random.default_rng(1)
    seen = set()
    for t in range(150):
        before = env.scalars()[L.S_SCENARIO].copy()      # the scenario every env is in while this step runs
        _, _, dones, infos = venv.step(rng.uniform(-4, 4, size=(B, 905)).astype(np.float32))
        for b in range(B):
            assert infos[b]["source_node"].startswith(f"s{before[b]}:"), (t, b, infos[b]["source_node"], before[b])
            assert infos[b]["target_node"].startswith(f"s{before[b]}:")
            seen.add(int(before[b]))
    assert len(seen) >= 3, "the device never switched scenarios: test too short"
    venv.close()


def test_raising_the_cutoffs_beyond_the_buffers_is_refused():
    """cbs_set_cutoffs after load must not let episodes outgrow the snapshot / edge buffers silently."""
    import ccbs_b200 as cb
    from ccbs_b200.batched_env import CbsError
    spec = cb.synthetic_spec(31, 20)
    env = cb.BatchedCyberBattleEnv([spec], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=4, seed=4)
    env.set_proportional_cutoff_coefficient(0.5)         # shorter episodes always fit
    with pytest.raises(CbsError, match="max_slots"):
        env.set_proportional_cutoff_coefficient(10)      # the reference's evaluation setting (utils/test_utils.py)
    env.close()
    big = cb.BatchedCyberBattleEnv([spec], cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=4, seed=4, max_slots=39, max_edges=200)
    big.set_proportional_cutoff_coefficient(10)          # sized for it at construction: accepted
    big.reset()
    rng = np.random.default_rng(0)
    import torch
    for _ in range(60):
        big.step(torch.from_numpy(rng.uniform(-4, 4, size=(4, 905)).astype(np.float32)).to(big.device), None)
    big.sync()                                           # no capacity flag
    big.close()
