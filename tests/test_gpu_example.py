"""-m gpu: the in-repo PPO example runs end to end on the batched env."""
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_ppo_example_runs(monkeypatch):
    import math
    sys.path.insert(0, "examples")
    import importlib
    mod = importlib.import_module("train_ppo_b200")
    monkeypatch.setattr(sys, "argv", ["train_ppo_b200.py", "--envs", "256", "--updates", "2", "--n-steps", "16", "--nodes", "10",
                                      "--scenarios", "3", "--minibatch", "1024", "--epochs", "2"])
    log = mod.main()
    assert len(log) == 2 and all(math.isfinite(r["ep_rew_mean"]) and r["episodes"] > 0 for r in log)
