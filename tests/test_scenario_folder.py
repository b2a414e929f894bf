"""Scenario folders in the reference's layout (pickled reference Models + split.yaml) load into the same tables as the
in-memory Models; reference-only (needs /root/reference for the Model class)."""
import os
import pickle

import numpy as np
import pytest
import yaml

import ccbs_b200 as cb

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")]


def test_load_scenario_folder_roundtrip(tmp_path):
    from oracle import ref_bridge as rb
    pool = cb.synthetic_vuln_pool(3, 40)
    models = {}
    for i in (1, 2, 3):
        g = cb.synthetic_input_graph(900 + i, 6 + i, pool=pool)
        models[i] = rb.reference_model_from_input_graph(g, seed=i)
        os.makedirs(tmp_path / str(i))
        with open(tmp_path / str(i) / "network_bert.pkl", "wb") as f:
            pickle.dump(models[i], f)
    with open(tmp_path / "split.yaml", "w") as f:
        yaml.safe_dump({"training_set": [{"id": 1}, {"id": 3}], "validation_set": [{"id": 2}]}, f)
    ids, specs = cb.load_scenario_folder(str(tmp_path), "bert", subset="training_set")
    assert ids == [1, 3]
    direct = cb.compile_scenarios([cb.spec_from_model(models[1]), cb.spec_from_model(models[3])])
    loaded = cb.compile_scenarios(specs)
    for name in ("row_packed", "vi_flags", "vi_success", "nd_ownable", "nd_discoverable", "recon_nodes", "outblock", "vemb64"):
        assert np.array_equal(getattr(direct, name), getattr(loaded, name)), name
    ids_all, _ = cb.load_scenario_folder(str(tmp_path), "bert")
    assert ids_all == [1, 2, 3]


def test_per_lm_feature_vectors_are_resolved():
    from oracle import ref_bridge as rb
    g = cb.synthetic_input_graph(950, 5, pool=cb.synthetic_vuln_pool(3, 40))
    model = rb.reference_model_from_input_graph(g, seed=1)
    want = cb.spec_from_model(model)
    for n in model.network.nodes:                      # un-resolve: {LM: vector} as the scenario generator stores them
        info = model.network.nodes[n]["data"]
        for s in info.services:
            s.feature_vector = {"bert": s.feature_vector, "gpt2": [0.0] * 768}
        for v in info.vulnerabilities.values():
            v.embedding = {"bert": v.embedding, "gpt2": [0.0] * 768}
    got = cb.spec_from_model(model, feature_extractor="bert")
    assert np.array_equal(cb.compile_scenarios([want]).vemb64, cb.compile_scenarios([got]).vemb64)
    with pytest.raises(ValueError):
        cb.spec_from_model(model)


def test_unsupported_reference_options_are_not_silently_dropped():
    rewards = {"rewards_dict": {"control": cb.config.DEFAULT_REWARDS["control"]},
               "penalties_dict": {"control": cb.config.DEFAULT_PENALTIES["control"]}}
    with pytest.raises(ValueError):
        cb.EnvConfig.from_reference_dicts({"static_defender_agent": "honeypot"}, rewards)
    # the re-imaging defender is implemented; its parameters come from the [min, max] ranges of train_config.yaml:39-44
    cfg = cb.EnvConfig.from_reference_dicts({"static_defender_agent": "reimage", "detect_probability_min": 0.05,
                                             "detect_probability_max": 0.15, "scan_capacity_min": 3, "scan_capacity_max": 3,
                                             "scan_frequency_min": 2, "scan_frequency_max": 4}, rewards)
    assert cfg.static_defender_agent == "reimage" and abs(cfg.detect_probability - 0.1) < 1e-12
    assert (cfg.scan_capacity, cfg.scan_frequency) == (3, 3)
    assert cb.EnvConfig.from_reference_dicts({"static_defender_agent": None}, rewards).static_defender_agent is None
    assert cb.EnvConfig.from_reference_dicts({"distance_metric": "l2"}, rewards).distance_metric == "l2"
    with pytest.raises(ValueError):                                    # compressed:578-579
        cb.EnvConfig.from_reference_dicts({"distance_metric": "chebyshev"}, rewards)
    # sample_subset_samples (compressed:553-567, the reference's training default) is carried through, not dropped, and the
    # caller's dict is left alone
    tc = {"sample_subset_samples": 100, "episode_iterations": 77, "static_defender_agent": None}
    before = dict(tc)
    cfg = cb.EnvConfig.from_reference_dicts(tc, rewards)
    assert cfg.episode_iterations == 77 and cfg.sample_subset_samples == 100 and tc == before
    assert cb.EnvConfig.from_reference_dicts({"static_defender_agent": "events"}, rewards).static_defender_agent == "events"
    assert cb.EnvConfig.from_reference_dicts({"switch_interval": 5}, rewards).switch_interval == 5
    assert cb.EnvConfig().switch_interval is None


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
def test_every_reference_train_config_parses():
    """EnvConfig.from_reference_dicts over the reference's own YAML files (agents/*/config/train_config.yaml +
    rewards_config.yaml) for all six goals: the defaults of every trainer are accepted, sample_subset_samples included."""
    import glob
    import warnings
    import yaml
    paths = glob.glob("/root/reference/cyberbattle/agents/**/config/train_config.yaml", recursive=True)
    assert len(paths) >= 3
    for tcp in paths:
        tc = yaml.safe_load(open(tcp))
        rcp = os.path.join(os.path.dirname(tcp), "rewards_config.yaml")
        rc = yaml.safe_load(open(rcp if os.path.exists(rcp) else "/root/reference/cyberbattle/agents/config/rewards_config.yaml"))
        assert set(rc["rewards_dict"]) == set(cb.constants.GOALS)
        for goal in rc["rewards_dict"]:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                cfg = cb.EnvConfig.from_reference_dicts(tc, rc, goal=goal)
            assert cfg.goal == goal and cfg.distance_metric == tc["distance_metric"]
            assert int(cfg.sample_subset_samples or 0) == int(tc.get("sample_subset_samples") or 0)
            assert cfg.switch_interval == tc.get("switch_interval")
            assert cfg.episode_iterations == tc["episode_iterations"]
            assert len(cfg.reward_vector()) == 10 and len(cfg.penalty_vector()) == 18
