"""-m gpu: BASELINE.json's full-size configurations.

configs[1] — 4096 envs of 10-25-node scenarios on one GPU: every 16th env (256 envs) is followed step by step by its own
oracle (bit-exact integer state, rewards / observations within 1e-5), ALL envs are checked against size-independent
invariants of the state (list lengths = popcounts, owned within discovered, order lists are permutations of the set bits, ...).
configs[2] — 65536 envs of 32-node scenarios (here on ONE GPU; the bench shards them 8192 per GPU): the same invariants on
all envs, and the first 96 envs reproduce bit for bit what a 96-env handle with the same global env indices computes
(results do not depend on the batch an env sits in)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _popcount(a):
    return np.unpackbits(np.ascontiguousarray(a).view(np.uint8), axis=-1).sum(axis=-1)


def check_invariants(env, num_nodes_of_env):
    """State invariants that hold for every env after any number of steps (no defender)."""
    from ccbs_b200 import constants as C, lib as L
    m, sc = env.masks(), env.scalars()           # [planes, words, B], [scalars, B]
    B = env.num_envs
    pc = lambda plane: _popcount(m[plane].T.copy()).reshape(B)  # noqa: E731
    owned, disc = m[C.M_OWNED], m[C.M_DISCOVERED]
    assert np.array_equal(pc(C.M_OWNED), sc[L.S_N_OWNED]), "n_owned != popcount(owned)"
    assert np.array_equal(pc(C.M_DISCOVERED), sc[L.S_N_DISC]), "n_disc != popcount(discovered)"
    assert not np.any(owned & ~disc), "an owned node is not discovered"
    assert not np.any(m[C.M_EXFILTRATED] & ~m[C.M_COLLECTED]), "exfiltrated without collected"
    assert not np.any(m[C.M_COLLECTED] & m[C.M_HAS_DATA]), "collected node still has data"
    assert not np.any(m[C.M_PRIV_ROOT] & ~m[C.M_PRIV_USER]), "root without user level"
    assert np.all(sc[L.S_N_DISC] <= num_nodes_of_env) and np.all(sc[L.S_N_OWNED] >= 1)
    assert np.array_equal(sc[L.S_STEPCOUNT], sc[L.S_NUM_ITER])
    assert np.array_equal(sc[L.S_SCST], (sc[L.S_SCENARIO] << 8) | sc[L.S_STARTER])
    # nothing beyond the scenario's node count is ever set
    for w in range(m.shape[1]):
        hi = np.clip(num_nodes_of_env - 32 * w, 0, 32)
        valid = np.where(hi >= 32, 0xFFFFFFFF, (1 << hi.astype(np.uint64)) - 1).astype(np.uint32)
        assert not np.any(m[:11, w, :] & ~valid[None, :]), "bit set beyond the scenario's nodes"
    # the order lists are permutations of the set bits (checked on a stride of envs: host loops)
    do, oo = env.disc_order(), env.owned_order()
    for b in range(0, B, max(1, B // 512)):
        nd, no = int(sc[L.S_N_DISC, b]), int(sc[L.S_N_OWNED, b])
        bits = lambda plane: {32 * w + i for w in range(m.shape[1]) for i in range(32) if (int(m[plane, w, b]) >> i) & 1}  # noqa: E731
        assert set(do[b, :nd].tolist()) == bits(C.M_DISCOVERED) and len(set(do[b, :nd].tolist())) == nd
        assert set(oo[b, :no].tolist()) == bits(C.M_OWNED) and len(set(oo[b, :no].tolist())) == no
        assert do[b, 0] == oo[b, 0] == sc[L.S_STARTER, b]


def test_config1_4096_envs_oracle_subsample():
    import torch
    import bench
    import ccbs_b200 as cb
    from ccbs_b200 import lib as L
    from oracle import trace as tr
    from oracle.cbs_oracle import OracleEnv
    from scipy.spatial import distance
    from tests.gpu_harness import masks_to_u64
    from tests.test_gpu_lockstep import philox_pick, philox_u01

    wl = bench.WORKLOADS["c1"]
    specs = bench.build_specs(wl)
    B, T, seed = wl["envs_per_gpu"], 40, 2027
    w, cfg = cb.GaeWeights.random(bench.GAE_SEED), cb.EnvConfig()
    env = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, seed=seed)
    sc_of_env = env.scenario_of_env
    nn = np.array([specs[s].num_nodes for s in sc_of_env])
    tables, g = env.tables, cb.constants.GOALS[cfg.goal]
    sample = list(range(0, B, 16))
    assert len(sample) == 256
    oracles = {b: OracleEnv(specs[sc_of_env[b]], w, cfg) for b in sample}
    vidx = {b: tr.vuln_index(specs[sc_of_env[b]]) for b in sample}
    episodes = {b: 0 for b in sample}

    def starter_for(b):
        sc = sc_of_env[b]
        f0, f1 = tables.sc_feasible_off[g][sc], tables.sc_feasible_off[g][sc + 1]
        return int(tables.feasible_starters[g][f0 + philox_pick(seed, b, episodes[b], 1, int(f1 - f0))])

    obs = env.reset()
    env.sync()
    obs_h = obs.cpu().numpy()
    for b in sample:
        o = oracles[b].reset(starter=starter_for(b))
        np.testing.assert_allclose(obs_h[b, :192], o["graph_embeddings"], rtol=1e-5, atol=2e-5)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(17)
    flips = 0
    for t in range(T):
        a = torch.rand(B, 905, device="cuda", generator=gen) * 8 - 4
        obs, reward, done, info = env.step(a, None)
        env.sync()
        a_h = a[sample].cpu().numpy()
        obs_h, rew_h, done_h, info_h = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), info.cpu().numpy()
        m = env.masks()
        for k, b in enumerate(sample):
            o = oracles[b]
            u = philox_u01(seed, b, t, 0)
            s, t_, vid, kind, d, _ = o.find_closest_action_embedding(a_h[k])
            want, got = (s, t_, vidx[b][vid], kind), tuple(int(x) for x in info_h[b, :4])
            forced = None
            if got != want:     # only a genuine near-tie of the oracle's own float64 distances is accepted
                dd = distance.cdist(np.atleast_2d(a_h[k]), o._rows_cache, "cosine").flatten()
                cand = [i for i, key in enumerate(o.action_keys) if (key[0], key[1], vidx[b][key[2]], key[3]) == got]
                assert cand, f"step {t} env {b}: decode {got} not in the oracle's table"
                i = min(cand, key=lambda j: dd[j])
                assert dd[i] - d < 1e-6, f"step {t} env {b}: decode {got} (gap {dd[i] - d:.3e}) vs oracle {want}"
                flips += 1
                forced = (got[0], got[1], o.action_keys[i][2], got[3], dd[i])
            ob, r, dn, _ = o.step(a_h[k], u, forced=forced)
            assert int(info_h[b, 4]) == o.outcome and int(info_h[b, 5]) == o.end_episode_reason and bool(done_h[b]) == bool(dn)
            np.testing.assert_allclose(rew_h[b], r, rtol=1e-5, atol=1e-4)
            if dn:
                episodes[b] += 1
                ob = o.reset(starter=starter_for(b))
            np.testing.assert_allclose(obs_h[b, :192], ob["graph_embeddings"], rtol=1e-5, atol=2e-5)
            assert np.array_equal(masks_to_u64(m, b), tr.masks_to_array(o.masks())), f"step {t} env {b}: masks differ"
        if t % 8 == 7 or t == T - 1:
            check_invariants(env, nn)
    assert flips <= 52, f"{flips} near-tie flips in {256 * T} env-steps"
    assert np.array_equal(env.scalars()[L.S_EPISODES][sample], np.array([episodes[b] for b in sample]))
    assert sum(episodes.values()) >= 256
    env.close()


def test_config2_65536_envs_invariants_and_batch_independence():
    import torch
    import bench
    import ccbs_b200 as cb
    wl = bench.WORKLOADS["c2"]
    specs = bench.build_specs(wl)
    B, T, K, seed = 65536, 36, 96, 7
    w, cfg = cb.GaeWeights.random(bench.GAE_SEED), cb.EnvConfig()
    env = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, seed=seed)
    small = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=K, seed=seed, scenario_of_env=env.scenario_of_env[:K])
    nn = np.array([specs[s].num_nodes for s in env.scenario_of_env])
    env.reset()
    small.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    for t in range(T):
        a = torch.rand(B, 905, device="cuda", generator=gen) * 8 - 4
        obs, rew, done, info = env.step(a, None)
        obs2, rew2, done2, info2 = small.step(a[:K].contiguous(), None)
        assert torch.equal(obs[:K], obs2) and torch.equal(rew[:K], rew2) and torch.equal(done[:K], done2) and torch.equal(info[:K], info2), f"step {t}"
        if t % 12 == 11:
            env.sync()
            check_invariants(env, nn)
    env.sync()
    small.sync()
    assert np.array_equal(env.masks()[:, :, :K], small.masks())
    acc = env.stat_accum()
    assert acc["episodes"] >= B and acc["cutoff"] + acc["wins"] + acc["lost"] == acc["episodes"]
    env.close()
    small.close()
