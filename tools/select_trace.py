"""Per-env cycle trace of decode_select (debug aid): which envs set the kernel's duration."""
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
trace = torch.zeros(B, 6, dtype=torch.int64, device="cuda")
env.lib.cbs_debug_select_trace(env._h, ct.c_void_p(trace.data_ptr()))
a = torch.rand(B, 905, device="cuda", generator=g) * 8 - 4
env.decode(a); env.sync()
t = trace.cpu().numpy()
cyc, rows, live, nex, combos, t0 = t.T
print("kernel span (cycles):", (t0 + cyc).max() - t0.min(), " = us at 1.9GHz:", ((t0 + cyc).max() - t0.min()) / 1900)
for name, v in (("cycles", cyc), ("rows", rows), ("live pairs", live), ("rescored", nex), ("combos", combos)):
    print(f"{name:11s} mean {v.mean():9.1f} p50 {np.percentile(v,50):8.0f} p90 {np.percentile(v,90):8.0f} p99 {np.percentile(v,99):8.0f} max {v.max():8.0f}")
order = np.argsort(-cyc)[:8]
print("slowest envs: cycles rows live rescored combos")
for b in order: print("   ", cyc[b], rows[b], live[b], nex[b], combos[b])
print("corr(cycles, rows) =", np.corrcoef(cyc, rows)[0,1], " corr(cycles, rescored) =", np.corrcoef(cyc, nex)[0,1], "corr(cycles, combos)=", np.corrcoef(cyc, combos)[0,1])
# start-time distribution: when did the last warp START?
print("last start at (us):", (t0.max() - t0.min()) / 1900, " mean start:", (t0.mean() - t0.min())/1900)
# least-squares model of the per-env cost: cycles ~ a + b * rows + c * ceil(combos / 32) + d * rescored
X = np.stack([np.ones(B), rows, np.ceil(combos / 32.0), nex], axis=1).astype(np.float64)
coef, *_ = np.linalg.lstsq(X, cyc.astype(np.float64), rcond=None)
print("fit: cycles = %.0f + %.2f * rows + %.0f * pair-blocks + %.0f * rescored" % tuple(coef))
print("shares of the mean: const %.0f%% rows %.0f%% blocks %.0f%% rescore %.0f%%" % tuple(100 * coef * X.mean(0) / cyc.mean()))
