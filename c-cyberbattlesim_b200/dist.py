"""Multi-GPU plumbing: the env batch shards over ranks with no communication on the step path; only the
episode-statistics accumulators are summed across ranks (one small all-reduce per logging interval).
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

from . import lib as L


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the global env index space owned by `rank`.  The start is also the
    handle's `global_env_offset`, which keys the Philox streams: results do not depend on the GPU count."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_episode_stats(accum: torch.Tensor, group=None) -> Dict[str, float]:
    """Sum the float64[20] accumulator vector (BatchedCyberBattleEnv.stat_accum_tensor()) over all ranks and
    return named totals plus the derived means SB3's logger reports."""
    t = accum.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = dict(zip(L.ACCUM_NAMES, t.cpu().tolist()))
    n = max(vals["episodes"], 1.0)
    vals["ep_rew_mean"] = vals["return_sum"] / n
    vals["ep_len_mean"] = vals["length_sum"] / n
    vals["win_rate"] = vals["wins"] / n
    return vals
