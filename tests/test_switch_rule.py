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


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
def test_node_detail_strings_match_get_str_info():
    """TraceCsvWriter._node_str (fed a state snapshot in the device's mask layout) against the reference's own
    RandomSwitchEnv.get_str_info (switch.py:307-334) for every node after every step of a scripted-attacker trace: status,
    privilege (IntEnum -> bare integer under the reference's Python 3.12), data / persistence / evasion flags, services with
    their firewall flags, vulnerabilities with the generator's outcome labels."""
    import ccbs_b200.constants as C
    from ccbs_b200 import lib as L
    from ccbs_b200.trace_csv import TraceCsvWriter
    from oracle import gen_golden as gg, trace as tr, ref_bridge as rb
    from oracle.cbs_oracle import OracleEnv
    name = "p6_control_win"
    p, graph, model, spec, cfg, weights = gg.build_case(name)
    case = gg.load_case(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    runner, oracle = rb.ReferenceRunner(model, weights, cfg), OracleEnv(spec, weights, cfg)

    class _Env:
        tables = cb.compile_scenarios([spec], cfg.isolation_filter_threshold)
    writer = TraceCsvWriter.__new__(TraceCsvWriter)
    writer.env = _Env()

    def snapshot():
        m64 = tr.masks_to_array(oracle.masks())                      # [N_MASKS, 2] uint64 -> the device's [N_MASKS, words, B] uint32
        masks = np.zeros((C.N_MASKS, 4, 1), np.uint32)
        for w in range(4):
            masks[:, w, 0] = (m64[:, w // 2] >> np.uint64(32 * (w % 2))) & np.uint64(0xFFFFFFFF)
        return dict(masks=masks, scal=np.zeros((L.NUM_SCALARS, 1), np.int32))

    starter = int(case["starters"][0])
    runner.reset(starter)
    oracle.reset(starter=starter)
    seen = set()
    for t in range(80):
        a = (np.asarray(oracle.action_rows[int(case["policy_rows"][t])], np.float64) + case["actions"][t].astype(np.float64)).astype(np.float32)
        runner.step(a, case["uniforms"][t])
        oracle.step(a, case["uniforms"][t])
        if oracle.done or oracle.truncated:
            break
        snap = snapshot()
        for j, nid in enumerate(runner.ids):
            want = runner.wrapper.get_str_info(runner.env.get_node(nid))
            assert writer._node_str(snap, 0, j) == want, (t, j)
            seen.add(want.split(" / tag")[0] + want.split("privilege level : ")[1][:1])
    assert t >= 40 and len(seen) >= 3          # Running / Stopped nodes and several privilege levels were compared
