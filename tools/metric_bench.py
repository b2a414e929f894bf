"""Step time of the fused cbs_step per decode metric (cosine / l1 / l2 / inf) on bench.py's configs[2] workload:
    python tools/metric_bench.py [envs] [steps]      -> one JSON line per metric (CUDA events on the launching stream)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
import ccbs_b200 as cb   # noqa: E402
from ccbs_b200 import constants as C   # noqa: E402
from ccbs_b200.batched_env import BatchedCyberBattleEnv   # noqa: E402
from ccbs_b200.gae import GaeWeights   # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    dev = torch.device("cuda", 0)
    specs = bench.build_specs(bench.WORKLOADS["c2"])
    weights = GaeWeights.random(bench.GAE_SEED)
    R = max(2, int(np.ceil(160e6 / (B * C.ACTION_DIM * 4))))     # action ring larger than the 126 MB L2
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    ring = (torch.rand(R, B, C.ACTION_DIM, device=dev, generator=gen) * 8.0 - 4.0).contiguous()
    for metric in ("cosine", "l1", "l2", "inf"):
        env = BatchedCyberBattleEnv(specs, weights, cb.EnvConfig(distance_metric=metric), num_envs=B, device=0, seed=7, auto_reset=True)
        env.reset()
        for i in range(10):
            env.step(ring[i % R], None, want_info=False)
        env.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            env.step(ring[(10 + i) % R], None, want_info=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        acc = env.stat_accum()
        print(json.dumps({"metric": metric, "envs": B, "steps": steps, "ms_per_step": round(ms, 4),
                          "env_steps_per_s": round(B / (ms * 1e-3)), "episodes": acc["episodes"],
                          "global_vulns": int(env.tables.vemb32.shape[0])}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
