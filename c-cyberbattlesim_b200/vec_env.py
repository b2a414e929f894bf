"""Stable-Baselines3-compatible surfaces over BatchedCyberBattleEnv.

* :class:`CyberBattleVecEnv` follows the SB3 ``VecEnv`` protocol (``num_envs``, spaces, ``reset``,
  ``step_async`` / ``step_wait`` with auto-reset, ``terminal_observation``, ``episode`` and
  ``TimeLimit.truncated`` info keys) and is the drop-in for the reference's
  ``DummyVecEnv([lambda: Monitor(RandomSwitchEnv(...))])`` / ``SubprocVecEnv`` construction sites
  (agents/train_agent.py:113, agents/multi_env/train_agent_multi_env.py:166-169).
* :class:`RandomSwitchEnvB200` is the single-env ``gymnasium.Env``-shaped surface of
  ``RandomSwitchEnv`` (_env/cyberbattle_env_switch.py:109-167,194-203).

``info`` carries the keys the reference's callbacks read (compressed:435-450, agents/callbacks.py:39-64,
agents/multi_env/callbacks_multi_env.py:20-32)."""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import constants as C
from . import spaces
from .batched_env import BatchedCyberBattleEnv

try:  # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:  # noqa: BLE001
    _SB3VecEnv = object


def _obs_dict(obs: np.ndarray):
    """[B, 194 | 258] flat observation -> the reference's Dict (the last two floats are the discrete features)."""
    return {"graph_embeddings": obs[:, :-2].astype(np.float64), "discrete_features": obs[:, -2:].astype(np.float64)}


class CyberBattleVecEnv(_SB3VecEnv):
    """``env``: a :class:`BatchedCyberBattleEnv`, or a :class:`~ccbs_b200.host_pipeline.ShardedHostEnv` (the same batch cut
    into handles whose host copies and kernels overlap — the faster choice when the trainer lives on the host)."""

    def __init__(self, env, lazy_infos: bool = False):
        self.env = env
        self.num_envs = env.num_envs
        self.observation_space = spaces.observation_space(env.obs_dim - 2)
        self.action_space = spaces.action_space()
        self.lazy_infos = lazy_infos
        self.render_mode = None
        self._actions = None
        B = self.num_envs
        self._h_actions = torch.empty(B, C.ACTION_DIM, dtype=torch.float32).pin_memory()
        self._h_obs = torch.empty(B, env.obs_dim, dtype=torch.float32).pin_memory()
        self._h_rew = torch.empty(B, dtype=torch.float32).pin_memory()
        self._h_done = torch.empty(B, dtype=torch.uint8).pin_memory()
        self._h_info = torch.empty(B, 8, dtype=torch.int32).pin_memory()
        self._ep_return = np.zeros(B, dtype=np.float64)
        self._ep_len = np.zeros(B, dtype=np.int64)
        self._t0 = time.time()
        self._empty = {}

    # ---- VecEnv protocol -------------------------------------------------------------------------
    def reset(self):
        obs = self.env.reset()
        self.env.sync()
        self._ep_return[:] = 0
        self._ep_len[:] = 0
        return _obs_dict(obs.cpu().numpy())

    def step_async(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, C.ACTION_DIM):
            raise ValueError(f"actions must have shape ({self.num_envs}, {C.ACTION_DIM})")
        self._h_actions.numpy()[...] = a

    def step_wait(self):
        env = self.env
        env.step_host(self._h_actions.numpy(), None, self._h_obs.numpy(), self._h_rew.numpy(), self._h_done.numpy(),
                      self._h_info.numpy())
        obs = self._h_obs.numpy().copy()
        rew = self._h_rew.numpy().astype(np.float32).copy()
        done = self._h_done.numpy().astype(bool)
        info = self._h_info.numpy()
        self._ep_return += rew
        self._ep_len += 1
        finished = np.nonzero(done)[0]
        term = stats = None
        if len(finished):
            term = env.terminal_obs()
            stats = env.last_stats()
        # info dicts: every env, or (lazy_infos) only the envs that finished — the others share one empty dict, so the
        # Python work per step is proportional to the number of episode ends, not to num_envs
        if self.lazy_infos:
            infos: List[dict] = [self._empty] * self.num_envs
            built = finished.tolist()
        else:
            infos = [None] * self.num_envs
            built = range(self.num_envs)
        if len(built):
            dist = env.distances().tolist()                  # one read per step, and only when some info dict is built
            rows = info.tolist()
            for b in built:
                d = self._info_dict(b, rows[b])
                d["min_distance_action"] = dist[b]           # compressed:449; callbacks_multi_env.py:89-90 reads it
                infos[b] = d
        for b in finished:
            d = infos[b]
            d["terminal_observation"] = {"graph_embeddings": term[b, :-2].astype(np.float64),
                                         "discrete_features": term[b, -2:].astype(np.float64)}
            # the reference reports done|truncated as `terminated` (compressed:451 via switch.py:121,148), so SB3 never
            # bootstraps at a cut-off; kept identical here
            d["TimeLimit.truncated"] = False
            d["truncated"] = bool(info[b, 7] & 1)
            d["episode"] = {"r": float(self._ep_return[b]), "l": int(self._ep_len[b]), "t": round(time.time() - self._t0, 6)}
            d["episode_stats"] = tuple(stats[b].tolist()[:13]) + (bool(stats[b, 13]),)   # callbacks_multi_env.py:20-32
            d["env_id"] = int(b)
            self._ep_return[b] = 0
            self._ep_len[b] = 0
        return _obs_dict(obs), rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _info_dict(self, b: int, row) -> dict:
        t = self.env.tables
        s, tg, u, kind, code, reason, step_count, last = (int(x) for x in row)
        sc = last >> 8      # the scenario in force during the step (the device switches scenarios on its own, switch.py:218-220)
        nodes, vulns = t.node_ids[sc], t.vuln_ids[sc]
        return {
            "description": "CyberBattleEnvCompressed step info",                               # compressed:436
            "source_node": nodes[s] if 0 <= s < len(nodes) else None,
            "target_node": nodes[tg] if 0 <= tg < len(nodes) else None,
            "source_node_tag": t.specs[sc].nodes[s].tag if t.specs and 0 <= s < len(nodes) else "",
            "target_node_tag": t.specs[sc].nodes[tg].tag if t.specs and 0 <= tg < len(nodes) else "",
            "vulnerability": vulns[u] if 0 <= u < len(vulns) else None,
            "vulnerability_type": "local" if s == tg else "remote",
            "outcome": C.KIND_INFO_STR[kind] if 0 <= kind < len(C.KIND_INFO_STR) else None,    # the DESIRED outcome (compressed:446-447)
            "outcome_class": C.KIND_NAMES[kind] if 0 <= kind < len(C.KIND_NAMES) else None,
            "outcome_obtained": C.KIND_NAMES[code] if code < 16 else C.OC_NAMES.get(code),
            "end_episode_reason": reason,
            "step_count": step_count,
        }

    def close(self):
        self.env.close()

    def seed(self, seed: Optional[int] = None):
        return [None] * self.num_envs     # randomness is the handle's Philox key (reference seed() is a no-op, cyberbattle_env.py:794)

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = self._indices(indices)
        if attr_name == "num_envs":
            return [1] * len(idx)
        if hasattr(self, attr_name):
            return [getattr(self, attr_name)] * len(idx)
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        idx = self._indices(indices)
        if method_name == "get_statistics":
            stats = self.env.last_stats()
            return [tuple(stats[b].tolist()[:13]) + (bool(stats[b, 13]),) for b in idx]
        if method_name == "set_cut_off":
            self.env.set_cut_off(*args)
            return [None] * len(idx)
        if method_name == "set_proportional_cutoff_coefficient":
            self.env.set_proportional_cutoff_coefficient(*args)
            return [None] * len(idx)
        raise AttributeError(f"env_method '{method_name}' is not available on the batched env")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(self._indices(indices))

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode: Optional[str] = None):
        return None

    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return list(indices)


class RandomSwitchEnvB200:
    """One env with the reference's gymnasium surface: reset() -> (obs, {}), step(a) -> (obs, reward, done,
    truncated, info), get_statistics(), set_cut_off(), set_proportional_cutoff_coefficient()
    (_env/cyberbattle_env_switch.py:109-167,194-203).  auto_reset is off: the caller resets, as gymnasium expects."""

    def __init__(self, specs, gae_weights, cfg=None, device: int = 0, switch_interval: Optional[int] = None, seed: int = 0,
                 interest_nodes=None):
        self.env = BatchedCyberBattleEnv(specs, gae_weights, cfg, num_envs=1, device=device, auto_reset=False,
                                         switch_interval=switch_interval, seed=seed, interest_nodes=interest_nodes)
        self.observation_space = spaces.observation_space(self.env.obs_dim - 2)
        self.action_space = spaces.action_space()
        self.num_envs = 1
        self.done = False
        self.truncated = False

    def reset(self, **kwargs):
        obs = self.env.reset()
        self.env.sync()
        o = _obs_dict(obs.cpu().numpy())
        self.done = False
        return {k: v[0] for k, v in o.items()}, {}

    def step(self, action):
        if self.done:
            raise RuntimeError("New episode must be started with env.reset()")     # cyberbattle_env.py:300-302
        a = torch.as_tensor(np.asarray(action, dtype=np.float32)).reshape(1, C.ACTION_DIM)
        obs, reward, done, info = self.env.step(a, None)
        self.env.sync()
        row = info.cpu().numpy()[0]
        self.done = bool(done.item())
        self.truncated = bool(row[7] & 1)
        o = _obs_dict(obs.cpu().numpy())
        d = CyberBattleVecEnv._info_dict(self, 0, row)
        d["min_distance_action"] = float(self.env.distances()[0])
        return {k: v[0] for k, v in o.items()}, float(self.env.reward64()[0]), self.done, self.truncated, d

    def get_statistics(self):
        s = self.env.last_stats()[0]
        return tuple(s.tolist()[:13]) + (bool(s[13]),)

    def set_cut_off(self, cut_off):
        self.env.set_cut_off(cut_off)

    def set_proportional_cutoff_coefficient(self, coefficient):
        self.env.set_proportional_cutoff_coefficient(coefficient)

    def close(self):
        self.env.close()
