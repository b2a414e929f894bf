"""-m gpu: cbs_replay — whole traces replayed on the device without a host round trip per step, compared with the reference's
recorded traces.  BASELINE configs[0] at its stated length: `g100_10k`, one default-like 100-node env, random actions, 10 000
steps, 112 episodes (the step-by-step harness of test_gpu_golden.py would need 10 000 host round trips for it)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _trace_from_log(log, e, n_nodes, first_obs, first_masks, defender=False):
    """cbs_replay's log of env `e` in the layout oracle/trace.py records"""
    import ccbs_b200.constants as C
    T = log["sel"].shape[0]

    def u64(m):                                         # uint32 [.., N_MASKS, words] -> uint64 [.., N_MASKS, 2]
        out = np.zeros(m.shape[:-1] + (2,), np.uint64)
        for w in range(m.shape[-1]):
            out[..., w // 2] |= m[..., w].astype(np.uint64) << np.uint64(32 * (w % 2))
        return out
    OW = 2 * n_nodes if defender else n_nodes
    rec = dict(sel=log["sel"][:, e], code=log["code"][:, e], reward=log["reward"][:, e], done=log["done"][:, e].astype(np.uint8),
               truncated=log["truncated"][:, e].astype(np.uint8), reason=log["reason"][:, e].astype(np.uint8), dist=log["dist"][:, e],
               masks=u64(log["masks"][:, e]), counters=log["counters"][:, e], obs=log["obs"][:, e], episode=log["episode"][:, e])
    disc = np.full((T, n_nodes), -1, np.int16)
    owned = np.full((T, OW), -1, np.int16)
    for t in range(T):
        nd, no = int(log["n_disc"][t, e]), int(log["n_owned"][t, e])
        disc[t, :nd] = log["disc_order"][t, e, :nd]
        owned[t, :no] = log["owned_order"][t, e, :no]
    rec["disc_order"], rec["owned_order"] = disc, owned
    fin = np.nonzero(rec["done"] | rec["truncated"])[0]
    rec["reset_obs"] = np.concatenate([first_obs[None], log["reset_obs"][fin, e]], axis=0).astype(np.float32)
    rec["reset_masks"] = np.concatenate([first_masks[None], u64(log["reset_masks"][fin, e])], axis=0)
    rec["stats"] = log["stats"][fin, e].reshape(-1, 14)
    rec["num_episodes"] = np.array(len(fin) + 1, np.int32)
    return rec


@pytest.mark.parametrize("name", ["g20_control", "g100_10k"])
def test_replay_reproduces_reference_trace(name, golden_dir):
    import torch
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from oracle import gen_golden as gg, trace as tr
    from tests.gpu_harness import masks_to_u64
    case = gg.load_case(os.path.join(golden_dir, name + ".npz"))
    B, T = 2, len(case["actions"])
    env = BatchedCyberBattleEnv([case["spec"]], case["weights"], case["cfg"], num_envs=B, auto_reset=True)
    env.set_starter_queue(np.tile(case["starters"][None, :], (B, 1)))
    actions = torch.from_numpy(np.ascontiguousarray(np.repeat(case["actions"][:, None, :], B, axis=1))).to(env.device)
    uniforms = torch.from_numpy(np.ascontiguousarray(np.repeat(case["uniforms"].astype(np.float32)[:, None], B, axis=1))).to(env.device)
    want = case["trace"]
    # A near-tie of two table rows (mathematically equal node embeddings that differ in the last float32 bit between the two
    # encoders) may decode to the other row.  The step-by-step harness follows the oracle there (TieFollower); a replay has no
    # host in the loop, so it is re-run with the RECORDED action forced at that step — accepted only if the device's own
    # float64 distance of its pick is within 1e-6 of the recorded one, and for fewer than 0.5 % of the steps.
    forced = {}
    for attempt in range(max(2, T // 200) + 1):
        obs0 = env.reset()
        env.sync()
        first_obs, first_masks = obs0[1].cpu().numpy().copy(), masks_to_u64(env.masks(), 1)
        l0 = env.launch_count
        log = env.replay(actions, uniforms, forced=forced)
        launches = env.launch_count - l0
        diff = np.nonzero(np.any(log["sel"][:, 1] != want["sel"], axis=1))[0]
        if len(diff) == 0:
            break
        t = int(diff[0])
        gap = abs(float(log["dist"][t, 1]) - float(want["dist"][t]))
        assert gap < 1e-6, f"step {t}: decoded {log['sel'][t, 1]} (d={log['dist'][t, 1]}) vs recorded {want['sel'][t]} (d={want['dist'][t]})"
        forced[t] = (want["sel"][t], want["dist"][t])
        # the starter queue is indexed by the env's lifetime episode count: a fresh handle starts the trace over
        env.close()
        env = BatchedCyberBattleEnv([case["spec"]], case["weights"], case["cfg"], num_envs=B, auto_reset=True)
        env.set_starter_queue(np.tile(case["starters"][None, :], (B, 1)))
    env.close()
    assert launches == 5 * T + 2 * len(forced)          # contraction, select + transition, observe, two log kernels (+ split steps)
    for k in ("sel", "code", "done", "masks", "disc_order", "n_disc"):       # identical inputs -> identical envs
        assert np.array_equal(log[k][:, 0], log[k][:, 1]), k
    rec = _trace_from_log(log, 1, case["spec"].num_nodes, first_obs, first_masks)
    report = tr.compare(rec, want, rtol=1e-5, atol=2e-5, label=f"replay/{name}")
    print(name, T, "steps,", int(rec["num_episodes"]), "episodes", report, "near-tie steps forced:", sorted(forced))
    assert int(rec["num_episodes"]) == int(want["num_episodes"]) and len(forced) <= max(1, T // 200)
