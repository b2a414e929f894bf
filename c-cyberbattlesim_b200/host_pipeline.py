"""ShardedHostEnv — one env batch on one GPU, held by K library handles and stepped through HOST buffers as a pipeline.

The end-to-end step of a host-side trainer (SB3 collects numpy actions, reference call site agents/train_agent.py:113)
is bound by the host-to-device copy of the [B, 905] float32 actions (29.7 MB at 8192 envs, ~0.55 ms over PCIe gen 5)
while the kernels of the whole batch take ~0.14 ms.  Episodes are independent, so the batch is cut into K contiguous
slices, each owned by its own handle and stream: slice k+1's copy-in runs under slice k's kernels and copy-out.  Philox
streams are keyed by the GLOBAL env index, so the results are those of one handle holding the whole batch
(tests/test_gpu_vecenv.py::test_sharded_host_env_matches_single_handle)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import constants as C
from .batched_env import BatchedCyberBattleEnv
from .config import EnvConfig
from .dist import shard_range
from .gae import GaeWeights, fold_gae
from .scenario import ScenarioSpec, compile_scenarios


def slice_bounds(num_envs: int, shards: int, weights: Optional[Sequence[float]] = None):
    """Contiguous [start, stop) env ranges of the pipeline's slices: equal (`shard_range`) or proportional to `weights`
    (every slice keeps at least one env)."""
    if weights is None or len(weights) != shards:
        return [shard_range(num_envs, k, shards) for k in range(shards)]
    w = np.asarray(weights, dtype=np.float64)
    if not np.all(w > 0):
        raise ValueError("shard_weights must be positive")
    cuts = np.round(np.cumsum(w) / w.sum() * num_envs).astype(np.int64)
    cuts[-1] = num_envs
    lo, out = 0, []
    for k, hi in enumerate(cuts):
        hi = int(min(max(hi, lo + 1), num_envs - (shards - 1 - k)))
        out.append((lo, hi))
        lo = hi
    return out


class ShardedHostEnv:
    def __init__(self, specs: Sequence[ScenarioSpec], gae_weights: GaeWeights, cfg: Optional[EnvConfig] = None,
                 num_envs: int = 1, shards: int = 4, device: int = 0, scenario_of_env: Optional[np.ndarray] = None,
                 seed: int = 0, global_env_offset: int = 0, interest_nodes: Optional[Sequence[int]] = None,
                 shard_weights: Optional[Sequence[float]] = None, **kw):
        """`shard_weights`: relative slice sizes (default: equal).  What a host step costs beyond the action copy is the LAST
        slice's kernels and copy-out, after the link has gone idle; the step-path kernels are latency bound, so a small last
        slice shortens that tail (bench.py's end-to-end leg uses tapering slices)."""
        cfg = cfg or EnvConfig()
        if shard_weights is not None:
            shards = len(shard_weights)
        shards = max(1, min(int(shards), int(num_envs)))
        node_goal = cfg.goal.endswith("node")
        tables = compile_scenarios(specs, cfg.isolation_filter_threshold, interest_nodes=interest_nodes if node_goal else None,
                                   interest_node_value=cfg.interest_node_value if node_goal else None)
        gae_tables = fold_gae(tables, gae_weights)
        if scenario_of_env is None:
            scenario_of_env = np.arange(num_envs, dtype=np.int32) % tables.num_scenarios
        self.scenario_of_env = np.ascontiguousarray(scenario_of_env, dtype=np.int32)
        self.bounds = slice_bounds(num_envs, shards, shard_weights)
        self.envs = [BatchedCyberBattleEnv(specs, gae_weights, cfg, num_envs=hi - lo, device=device,
                                           scenario_of_env=self.scenario_of_env[lo:hi], seed=seed,
                                           global_env_offset=global_env_offset + lo, tables=tables, gae_tables=gae_tables, **kw)
                     for lo, hi in self.bounds]
        self.cfg, self.tables, self.num_envs, self.obs_dim = cfg, tables, int(num_envs), self.envs[0].obs_dim
        self.device = self.envs[0].device

    def reset(self) -> torch.Tensor:
        return torch.cat([e.reset() for e in self.envs], dim=0)

    def sync(self):
        for e in self.envs:
            e.sync()

    def close(self):
        for e in self.envs:
            e.close()

    def step_host(self, actions: np.ndarray, uniforms: Optional[np.ndarray], obs: np.ndarray, reward: np.ndarray,
                  done: np.ndarray, info: Optional[np.ndarray] = None):
        """Same contract as :meth:`BatchedCyberBattleEnv.step_host` for the whole batch (row slices of the caller's
        arrays are handed to the shards in place)."""
        assert actions.shape == (self.num_envs, C.ACTION_DIM)
        for (lo, hi), e in zip(self.bounds, self.envs):
            e.step_host_async(actions[lo:hi], None if uniforms is None else uniforms[lo:hi], obs[lo:hi], reward[lo:hi],
                              done[lo:hi], None if info is None else info[lo:hi])
        for e in self.envs:
            e.host_sync()

    def set_cut_off(self, cut_off: int):
        for e in self.envs:
            e.set_cut_off(cut_off)

    def set_proportional_cutoff_coefficient(self, coefficient: float):
        for e in self.envs:
            e.set_proportional_cutoff_coefficient(coefficient)

    # ---- what CyberBattleVecEnv reads at episode ends ----
    def terminal_obs(self) -> np.ndarray:
        return np.concatenate([e.terminal_obs() for e in self.envs], axis=0)

    def last_stats(self) -> np.ndarray:
        return np.concatenate([e.last_stats() for e in self.envs], axis=0)

    def distances(self) -> np.ndarray:
        return np.concatenate([e.distances() for e in self.envs], axis=0)

    def stat_accum(self) -> dict:
        out: dict = {}
        for e in self.envs:
            for k, v in e.stat_accum().items():
                out[k] = out.get(k, 0.0) + v
        return out

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.envs)

    @property
    def state_bytes(self) -> int:
        return sum(e.state_bytes for e in self.envs)
