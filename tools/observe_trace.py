"""Per-item timeline of the observe kernel (debug aid): which items set the kernel's duration."""
import os, sys, ctypes as ct
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ccbs_b200 as cb
from ccbs_b200.batched_env import BatchedCyberBattleEnv
B = 8192
env = BatchedCyberBattleEnv(bench.build_specs(bench.WORKLOADS["c2"]), cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=7)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
for i in range(150):
    env.step(torch.rand(B, 905, device="cuda", generator=g) * 8 - 4, None, want_info=False)
trace = torch.zeros(B, 9, dtype=torch.int64, device="cuda")
env.lib.cbs_debug_observe_trace(env._h, ct.c_void_p(trace.data_ptr()))
env.step(torch.rand(B, 905, device="cuda", generator=g) * 8 - 4, None, want_info=False); env.sync()
env.lib.cbs_debug_observe_trace(env._h, None)
t = trace.cpu().numpy()
t = t[t[:, 0] > 0]
start, dur, flags, ne = t.T[:4]
ph = t[:, 4:]
t0 = start.min()
print("items", len(t), "span us", (start + dur).max() / 1e3 - t0 / 1e3, " last start us", (start.max() - t0) / 1e3)
cls = np.where(flags & 128, 0, np.where(flags & 32, 1, 2))
for c, name in enumerate(("episode end", "re-encode", "edge only")):
    m = cls == c
    if m.any():
        print(f"{name:12s} n {m.sum():5d}  dur us mean {dur[m].mean() / 1e3:6.2f} p50 {np.percentile(dur[m], 50) / 1e3:6.2f} p99 {np.percentile(dur[m], 99) / 1e3:6.2f} max {dur[m].max() / 1e3:6.2f}"
              f"  start us mean {(start[m] - t0).mean() / 1e3:6.2f} max {(start[m] - t0).max() / 1e3:6.2f}  nodes mean {(ne[m] >> 16).mean():.1f} edges mean {(ne[m] & 0xFFFF).mean():.1f}")
print("sum of durations / 1184 warps (us):", dur.sum() / 1184 / 1e3)
end = start + dur - t0
print("finish percentiles us:", [round(float(np.percentile(end, q)) / 1e3, 1) for q in (10, 50, 90, 99, 100)])
m = cls == 1
print("corr(dur, nodes)", np.corrcoef(dur[m], ne[m] >> 16)[0, 1], "corr(dur, edges)", np.corrcoef(dur[m], ne[m] & 0xFFFF)[0, 1])

m = cls == 1
print("re-encode items: encode done at", ph[m, 0].mean() / 1e3, "table done at", ph[m, 1].mean() / 1e3, "total", dur[m].mean() / 1e3, "(us; encode time includes the edge update)")
m = cls == 0
enc_first = np.where(ph[m, 0] > 0, ph[m, 0], 0)
print("episode-end items: [re-encode done", ph[m, 0].mean() / 1e3, "table", ph[m, 1].mean() / 1e3, "] finish done", ph[m, 2].mean() / 1e3, "reset done", ph[m, 3].mean() / 1e3,
      "encode done", ph[m, 4].mean() / 1e3, "total", dur[m].mean() / 1e3, " share with re-encode first:", (ph[m, 0] > 0).mean())
