"""The l1 / l2 / inf decode (csrc/k_decode_metric.cu) never forms a table row: it joins four per-part norms — source embedding,
target embedding, vulnerability embedding, outcome one-hot.  Host restatement of that join against np.linalg.norm over the
oracle's explicit rows (compressed:571-576), on a table grown by a scripted attacker."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("metric,ord_", [("l1", 1), ("l2", 2), ("inf", np.inf)])
def test_row_norms_separate_over_the_four_parts(metric, ord_, golden_dir):
    from oracle import gen_golden as gg
    from oracle.cbs_oracle import OracleEnv
    case = gg.load_case(os.path.join(golden_dir, "p8_l2.npz"))
    cfg = case["cfg"]
    cfg.distance_metric = metric
    env = OracleEnv(case["spec"], case["weights"], cfg)
    env.reset(starter=int(case["starters"][0]))
    for t in range(60):                      # grow the table: replay the recorded rows of the scripted attacker
        a = (np.asarray(env.action_rows[int(case["policy_rows"][t])], np.float64) + case["actions"][t].astype(np.float64)).astype(np.float32)
        env.step(a, case["uniforms"][t])
        if env.done or env.truncated:
            break
    rows = np.array(env.action_rows)
    assert rows.shape[0] > 200 and rows.shape[1] == 905
    action = np.random.default_rng(0).uniform(-4, 4, 905).astype(np.float32)
    want = env.all_distances(action)
    assert np.array_equal(want, np.linalg.norm(np.atleast_2d(action) - rows, ord=ord_, axis=1))
    diff = np.abs(action.astype(np.float64)[None, :] - rows)
    cuts = [(0, 64), (64, 128), (128, 896), (896, 905)]
    if metric == "l1":
        got = sum(diff[:, a:b].sum(1) for a, b in cuts)
    elif metric == "l2":
        got = np.sqrt(sum((diff[:, a:b] ** 2).sum(1) for a, b in cuts))
    else:
        got = np.max(np.stack([diff[:, a:b].max(1) for a, b in cuts]), axis=0)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    assert int(np.argmin(got)) == int(np.argmin(want))
    # the last nine columns are exactly one-hot, which is what lets the kernel tabulate that part per outcome column
    onehot = rows[:, 896:]
    assert np.all((onehot == 0) | (onehot == 1)) and np.all(onehot.sum(1) == 1)
