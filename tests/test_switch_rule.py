"""S1: when the switcher draws a new scenario.  The reference's own RandomSwitchEnv over three compressed envs, with
np.random.choice answering from a pre-drawn list, against ccbs_b200.constants.switch_due — the test the device applies when it
resets a finished env in place (tests/test_gpu_sharding.py::test_scenario_switch_cadence checks the device against the same
function)."""
import os

import numpy as np
import pytest

import ccbs_b200 as cb
from ccbs_b200.constants import switch_due


def test_switch_due_values():
    assert [switch_due(e, 2) for e in range(7)] == [False, False, True, False, False, True, False]
    assert all(switch_due(e, 0) for e in range(4))                       # interval 0: every reset
    assert [switch_due(e, 5) for e in (4, 5, 10, 11)] == [False, True, False, True]


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
@pytest.mark.parametrize("interval", [1, 2, 4])
def test_reference_switcher_follows_switch_due(interval):
    from oracle import ref_bridge as rb
    pool = cb.synthetic_vuln_pool(7, 40)
    models = []
    for k in range(3):
        graph = cb.synthetic_input_graph(300 + k, 5 + k, pool=pool)
        models.append(rb.reference_model_from_input_graph(graph, seed=300 + k))
    cfg = cb.EnvConfig(goal="control", proportional_cutoff_coefficient=1, episode_iterations=12, isolation_filter_threshold=0.0)
    picks = (np.arange(64) * 2 + interval) % 3          # every draw changes the scenario
    n_episodes = 11
    seq, used = rb.reference_switch_sequence(models, cb.GaeWeights.random(0), cfg, interval, picks, n_episodes)
    # restatement: the constructor draws once (switch.py:37); reset() number e (after e finished episodes) draws iff switch_due
    want, cur, i = [], int(picks[0]), 1
    for e in range(n_episodes):
        if switch_due(e, interval):
            cur, i = int(picks[i]), i + 1
        want.append(cur)
    assert seq == want and used == i
    assert len(set(seq)) > 1


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
def test_csv_header_is_the_reference_header(tmp_path):
    """ccbs_b200.trace_csv.HEADER against the header row the reference's own switcher writes (switch.py:223-243, the variant
    without embeddings)."""
    import csv
    from ccbs_b200.trace_csv import HEADER
    from oracle import ref_bridge as rb
    graph = cb.synthetic_input_graph(310, 5, pool=cb.synthetic_vuln_pool(7, 40))
    model = rb.reference_model_from_input_graph(graph, seed=310)
    runner = rb.ReferenceRunner(model, cb.GaeWeights.random(0), cb.EnvConfig(isolation_filter_threshold=0.0))
    runner.fake.next_starter = 0
    wrapper = runner.ref["switch"].RandomSwitchEnv(envs_ids=[0], switch_interval=10 ** 9, envs_list=[runner.env], verbose=0,
                                                   save_to_csv=True, csv_folder=str(tmp_path), save_embeddings=False)
    wrapper.file.flush()
    with open(os.path.join(str(tmp_path), "logs.csv"), newline="") as f:
        assert next(csv.reader(f)) == HEADER
    wrapper.file.close()
