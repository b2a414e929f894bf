"""-m gpu: results do not depend on how the env batch is sharded (Philox streams are keyed by the GLOBAL env index),
and on-device scenario switching follows RandomSwitchEnv._check_switch (cyberbattle_env_switch.py:218-220)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(env, actions, steps):
    import torch
    env.reset()
    out = []
    for t in range(steps):
        obs, rew, done, info = env.step(torch.from_numpy(actions[t]).to(env.device), None)
        env.sync()
        out.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy(), info.cpu().numpy().copy()))
    return out, env.masks(), env.scalars()


def test_two_shards_equal_one_batch():
    import ccbs_b200 as cb
    from ccbs_b200 import dist as cd
    specs = [cb.synthetic_spec(700 + k, 12) for k in range(4)]
    w, cfg = cb.GaeWeights.random(1), cb.EnvConfig()
    B, T = 48, 50
    sc = (np.arange(B) % 4).astype(np.int32)
    rng = np.random.default_rng(3)
    actions = rng.uniform(-4, 4, size=(T, B, 905)).astype(np.float32)
    full = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, scenario_of_env=sc, seed=99)
    ref, ref_masks, ref_sc = _run(full, actions, T)
    full.close()
    for world in (2, 3):
        for rank in range(world):
            lo, hi = cd.shard_range(B, rank, world)
            shard = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=hi - lo, scenario_of_env=sc[lo:hi], seed=99,
                                             global_env_offset=lo)
            got, masks, scal = _run(shard, np.ascontiguousarray(actions[:, lo:hi]), T)
            shard.close()
            for t in range(T):
                for a, b in zip(got[t], ref[t]):
                    assert np.array_equal(a, b[lo:hi]), f"world {world} rank {rank} step {t}"
            assert np.array_equal(masks, ref_masks[:, :, lo:hi])
            from ccbs_b200 import lib as L
            keep = [k for k in range(L.NUM_SCALARS) if k not in (L.S_TOTAL_STEPS, L.S_N_ENCODES)]
            assert np.array_equal(scal[keep], ref_sc[keep][:, lo:hi])


def test_scenario_switch_cadence():
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    from tests.test_gpu_lockstep import philox_pick
    specs = [cb.synthetic_spec(800 + k, 8 + k) for k in range(5)]
    B, seed, interval = 16, 5, 2
    env = cb.BatchedCyberBattleEnv(specs, cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=seed,
                                   switch_interval=interval, scenario_of_env=np.zeros(B, np.int32))
    env.reset()
    env.sync()
    want = np.zeros(B, np.int64)
    episodes = np.zeros(B, np.int64)
    rng = np.random.default_rng(0)
    for t in range(120):
        a = torch.from_numpy(rng.uniform(-4, 4, size=(B, 905)).astype(np.float32)).to(env.device)
        _, _, done, _ = env.step(a, None)
        env.sync()
        d = done.cpu().numpy().astype(bool)
        for b in np.nonzero(d)[0]:
            episodes[b] += 1
            # _check_switch: (episode_count + 1) % (switch_interval + 1) == 0 -> new scenario (uniform choice)
            if cb.constants.switch_due(episodes[b], interval):
                want[b] = philox_pick(seed, b, int(episodes[b]), 2, len(specs))
        sc = env.scalars()
        assert np.array_equal(sc[L.S_SCENARIO], want), f"step {t}"
        assert np.array_equal(sc[L.S_EPISODES], episodes)
        # the observation's discrete features are consistent with the scenario actually loaded
        assert np.all(sc[L.S_N_DISC] <= np.array([specs[s].num_nodes for s in sc[L.S_SCENARIO]]))
    assert len(set(want.tolist())) > 1 and episodes.min() >= 3
    env.close()
