"""CyberBattleVecEnv.step_wait host logic without a GPU: a stand-in for the batched env fills the host buffers the way
cbs_step_host does; the info dicts must carry the keys of the reference's StepInfo (compressed:41-57) that its callbacks read
(agents/multi_env/callbacks_multi_env.py:20-32,74-128), SB3's episode keys at episode ends, and nothing for running envs under
lazy_infos."""
import os

import numpy as np
import pytest
import torch

import ccbs_b200 as cb
import ccbs_b200.constants as C


class _FakeBatchedEnv:
    """The attributes / methods CyberBattleVecEnv touches (vec_env.py), host only."""

    def __init__(self, specs, B):
        self.tables = cb.compile_scenarios(specs, 0.1)
        self.num_envs, self.obs_dim = B, C.OBS_DIM + 2
        self.scenario_of_env = np.arange(B, dtype=np.int32) % len(specs)
        self.t = 0

    def step_host(self, actions, uniforms, obs, reward, done, info):
        assert actions.shape == (self.num_envs, C.ACTION_DIM) and actions.dtype == np.float32
        self.t += 1
        obs[...] = np.arange(obs.size, dtype=np.float32).reshape(obs.shape) * 1e-3
        reward[...] = np.arange(self.num_envs) - 1.5
        done[...] = [(b + self.t) % 3 == 0 for b in range(self.num_envs)]
        for b in range(self.num_envs):
            # source, target, local vulnerability, desired kind, obtained code, end reason, step count, truncated | scenario << 8
            info[b] = [0, 1 % self.tables.specs[self.scenario_of_env[b]].num_nodes, 0, C.K_RECON, C.OC_REPEATED,
                       3 if done[b] else 0, self.t, int(done[b]) | (int(self.scenario_of_env[b]) << 8)]

    def distances(self):
        return 0.25 + np.arange(self.num_envs, dtype=np.float64)

    def terminal_obs(self):
        return np.full((self.num_envs, self.obs_dim), 7.0, np.float32)

    def last_stats(self):
        s = np.zeros((self.num_envs, 14))
        s[:, 0], s[:, 13] = 2, 1
        return s

    def close(self):
        pass


@pytest.fixture
def no_pinning(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)


@pytest.mark.parametrize("lazy", [False, True])
def test_step_wait_info_dicts(no_pinning, lazy):
    from ccbs_b200.vec_env import CyberBattleVecEnv
    specs = [cb.synthetic_spec(40 + k, 6 + k) for k in range(2)]
    B = 5
    venv = CyberBattleVecEnv(_FakeBatchedEnv(specs, B), lazy_infos=lazy)
    for step in range(1, 4):
        obs, rew, done, infos = venv.step(np.zeros((B, C.ACTION_DIM), np.float32))
        assert obs["graph_embeddings"].shape == (B, C.OBS_DIM) and obs["discrete_features"].shape == (B, 2)
        assert rew.dtype == np.float32 and done.dtype == bool and len(infos) == B
        for b in range(B):
            d = infos[b]
            if lazy and not done[b]:
                assert d == {}
                continue
            sc = b % 2
            assert d["source_node"] == venv.env.tables.node_ids[sc][0] and d["target_node"] == venv.env.tables.node_ids[sc][1]
            assert d["vulnerability"] == venv.env.tables.vuln_ids[sc][0] and d["vulnerability_type"] == "remote"
            assert d["outcome"] == "Reconnaissance" and d["step_count"] == step and d["min_distance_action"] == 0.25 + b
            assert d["source_node_tag"] == specs[sc].nodes[0].tag
            if done[b]:
                assert d["end_episode_reason"] == 3 and d["TimeLimit.truncated"] is False and d["truncated"] is True
                assert set(d["episode"]) == {"r", "l", "t"} and d["env_id"] == b and len(d["episode_stats"]) == 14
                assert d["episode_stats"][13] is True and d["terminal_observation"]["graph_embeddings"].shape == (C.OBS_DIM,)
            else:
                assert "episode" not in d and d["end_episode_reason"] == 0
    venv.close()


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/cyberbattle"), reason="reference tree not mounted")
def test_info_keys_cover_the_reference_stepinfo(no_pinning):
    from ccbs_b200.vec_env import CyberBattleVecEnv
    from oracle import ref_bridge as rb
    step_info = rb.import_reference()["compressed"].StepInfo
    venv = CyberBattleVecEnv(_FakeBatchedEnv([cb.synthetic_spec(40, 6)], 2))
    _, _, _, infos = venv.step(np.zeros((2, C.ACTION_DIM), np.float32))
    # duration_in_ms is wall-clock bookkeeping of the Python env; network_availability travels in episode_stats[8]
    assert set(step_info.__annotations__) - set(infos[0]) == {"duration_in_ms", "network_availability"}
