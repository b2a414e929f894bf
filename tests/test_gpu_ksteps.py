"""-m gpu: cbs_transition_ksteps (K transitions per env in one launch, records in registers) against K cbs_transition launches on
the same pre-decoded actions: rewards, done flags, mask records, list lengths and order lists bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("uniforms_given", [True, False], ids=["uniforms", "philox"])
def test_ksteps_equals_k_single_transitions(uniforms_given):
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import constants as C, lib as L
    specs = [cb.synthetic_spec(300 + k, 10 + 3 * k) for k in range(6)]
    B, K, seed = 96, 12, 17
    cfg = cb.EnvConfig(proportional_cutoff_coefficient=0.5)
    w = cb.GaeWeights.random(0)
    env = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, seed=seed, auto_reset=False)
    t = env.tables
    g = C.GOALS["control"]
    first = np.asarray(t.feasible_starters[g])[np.asarray(t.sc_feasible_off[g])[:-1]]
    env.set_starter_queue(first[env.scenario_of_env][:, None].astype(np.int32))
    gen = torch.Generator(device=env.device)
    gen.manual_seed(5)
    sel_all = torch.empty(K, B, 4, dtype=torch.int32, device=env.device)
    dist_all = torch.empty(K, B, dtype=torch.float64, device=env.device)
    uni = torch.rand(K, B, device=env.device, generator=gen) if uniforms_given else None
    rew, don = [], []
    env.reset()
    for k in range(K):
        a = torch.rand(B, 905, device=env.device, generator=gen) * 8 - 4
        sel, dist = env.decode(a)
        sel_all[k].copy_(sel)
        dist_all[k].copy_(dist)
        r, d, _, _ = env.transition(sel, dist, None if uni is None else uni[k])
        rew.append(r.clone())
        don.append(d.clone())
        env.observe()
    env.sync()
    want = dict(masks=env.masks(), scal=env.scalars(), disc=env.disc_order(), owned=env.owned_order())
    assert sum(int(d.sum()) for d in don) > 0, "no env finished: the finished-env branch is not exercised"
    if not uniforms_given:       # Philox draws are keyed by the env's lifetime step counter: start the replay from the same count
        env.close()
        env = cb.BatchedCyberBattleEnv(specs, w, cfg, num_envs=B, seed=seed, auto_reset=False)
        env.set_starter_queue(first[env.scenario_of_env][:, None].astype(np.int32))
    env.reset()
    rk, dk = env.transition_ksteps(sel_all, dist_all, uni)
    env.sync()
    for k in range(K):
        assert torch.equal(rk[k], rew[k]) and torch.equal(dk[k], don[k]), f"step {k}"
    got = dict(masks=env.masks(), scal=env.scalars(), disc=env.disc_order(), owned=env.owned_order())
    assert np.array_equal(got["masks"], want["masks"])
    for name in ("S_FLAGS", "S_STEPCOUNT", "S_NUM_ITER", "S_OUTCOME", "S_N_DISC", "S_N_OWNED", "S_DISC_AMOUNT", "S_EP_RETURN", "S_EP_RETURN_HI"):
        k = getattr(L, name)
        a, b = got["scal"][k], want["scal"][k]
        if name == "S_FLAGS":        # the per-step work flags (add edge / re-encode / dirty) are consumed by observe in the stepwise run
            a, b = a & 0xF, b & 0xF
        assert np.array_equal(a, b), name
    nd, no = want["scal"][L.S_N_DISC], want["scal"][L.S_N_OWNED]
    for b in range(B):
        assert np.array_equal(got["disc"][b, :nd[b]], want["disc"][b, :nd[b]])
        assert np.array_equal(got["owned"][b, :no[b]], want["owned"][b, :no[b]])
    env.close()
