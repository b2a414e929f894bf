"""-m gpu: SURVEY §8a hazard 11 — an all-zero action makes every cosine distance NaN (0 / 0 in scipy's cdist) and np.argmin
returns the FIRST NaN row, i.e. the first row of the insertion-ordered table; the distance penalty then makes the reward NaN.
decode_select keeps np.argmin's NaN rule in its float64 re-score (k_decode.cu flush_candidates); the l1 / l2 / inf kernels
never produce NaN from finite inputs and must simply agree with the oracle.

Distances are compared at the golden replay's float tolerance (rtol 1e-5: the node embeddings come from two different fp32
encoders, torch-CPU in the oracle and the observe kernel here); the decoded row, the NaN-ness and "row 0 wins" are exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_zero_action_decodes_like_np_argmin(metric):
    import warnings
    import torch
    import ccbs_b200 as cb
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from oracle import trace as tr
    from oracle.cbs_oracle import OracleEnv
    spec, w, cfg = cb.synthetic_spec(21, 12), cb.GaeWeights.random(0), cb.EnvConfig(distance_metric=metric)
    oracle = OracleEnv(spec, w, cfg)
    starter = oracle.feasible_starters()[0]
    oracle.reset(starter=starter)
    env = BatchedCyberBattleEnv([spec], w, cfg, num_envs=3, auto_reset=True)
    env.set_starter_queue(np.full((3, 4), starter, np.int32))
    env.reset()
    vidx = tr.vuln_index(spec)
    rng = np.random.default_rng(0)
    for t in range(12):                                   # grow the table a little, then send zeros
        a = np.zeros(905, np.float32) if t in (0, 6, 11) else rng.uniform(-4, 4, 905).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s, t_, vid, kind, d, row = oracle.find_closest_action_embedding(a)
        sel, dist = env.decode(torch.from_numpy(a).to(env.device).unsqueeze(0).repeat(3, 1).contiguous())
        env.sync()
        got, gd = sel.cpu().numpy(), dist.cpu().numpy()
        assert all(tuple(int(x) for x in got[b]) == (s, t_, vidx[vid], kind) for b in range(3)), (t, got, (s, t_, vidx[vid], kind))
        if np.isnan(d):
            assert metric == "cosine" and row == 0 and np.all(np.isnan(gd))
        else:
            np.testing.assert_allclose(gd, d, rtol=1e-5)
        u = torch.full((3,), 0.37, dtype=torch.float32, device=env.device)
        env.transition(sel, dist, u)
        env.observe()
        env.sync()
        oracle.step(a, 0.37)
        if oracle.done or oracle.truncated:
            break
        if np.isnan(d):
            assert np.all(np.isnan(env.reward64())) and np.isnan(oracle.reward)
    env.close()
