"""Per-item timeline of the observe kernel (debug aid): which items set the kernel's duration."""
import os, sys, ctypes as ct
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ccbs_b200 as cb
from ccbs_b200.batched_env import BatchedCyberBattleEnv
B = int(os.environ.get("CBS_TRACE_B", "8192"))
env = BatchedCyberBattleEnv(bench.build_specs(bench.WORKLOADS["c2"]), cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=7)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
stagger = torch.arange(B, device="cuda") % 32   # episode phases spread like bench.py's pre-steps
for i in range(150):
    if i < 32:
        env.reset((stagger == i).to(torch.uint8))
    env.step(torch.rand(B, 905, device="cuda", generator=g) * 8 - 4, None, want_info=False)
TRW = 18 if os.environ.get("CBS_OBS_SUBTRACE") else 9
trace = torch.zeros(B, TRW, dtype=torch.int64, device="cuda")
env.lib.cbs_debug_observe_trace(env._h, ct.c_void_p(trace.data_ptr()))
env.step(torch.rand(B, 905, device="cuda", generator=g) * 8 - 4, None, want_info=False); env.sync()
env.lib.cbs_debug_observe_trace(env._h, None)
t = trace.cpu().numpy()
t = t[t[:, 0] > 0]
start, dur, flags, ne = t.T[:4]
wid = flags >> 32
flags = flags & 0xFFFFFFFF
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

# per-warp view: who sets the kernel's duration
nw = int(wid.max()) + 1
busy = np.bincount(wid, weights=dur, minlength=nw) / 1e3
last = np.zeros(nw); np.maximum.at(last, wid, (start + dur - t0) / 1e3)
first = np.full(nw, 1e9); np.minimum.at(first, wid, (start - t0) / 1e3)
print("warps", nw, "busy us mean", busy.mean(), "max", busy.max(), " last finish p50/p90/max", np.percentile(last, 50), np.percentile(last, 90), last.max(),
      " first start min/max", first.min(), first[first < 1e9].max())
order = np.argsort(-(start + dur))[:12]
for k in order:
    print("late item: warp", wid[k], "class", cls[k], "start", (start[k] - t0) / 1e3, "dur", dur[k] / 1e3, "nodes", ne[k] >> 16, "edges", ne[k] & 0xFFFF, "phases", (ph[k] / 1e3).round(1))
order = np.argsort(-dur)[:12]
for k in order:
    print("long item: warp", wid[k], "class", cls[k], "start", (start[k] - t0) / 1e3, "dur", dur[k] / 1e3, "nodes", ne[k] >> 16, "edges", ne[k] & 0xFFFF, "phases", (ph[k] / 1e3).round(1))
m = cls == 1
nn = ne[m] >> 16
for lo, hi in ((1, 4), (5, 8), (9, 12), (13, 16), (17, 24), (25, 32)):
    q = (nn >= lo) & (nn <= hi)
    if q.any(): print(f"re-encode nodes {lo}-{hi}: n {q.sum()} dur mean {dur[m][q].mean() / 1e3:.1f} max {dur[m][q].max() / 1e3:.1f} start mean {(start[m][q] - t0).mean() / 1e3:.1f}")

if os.environ.get("CBS_OBS_SUBTRACE"):   # library built with -DCBS_OBS_SUBTRACE: 14 phase slots; 5.. = the encode's sub-phases, 13 = edge update done
    m = (cls == 1) & (ph[:, 5] > 0) & (ph[:, 4] > 0)
    nn = ne[m] >> 16
    names = ["edge_upd", "scalars", "pack", "root", "nnconv", "bn1", "proj", "agg", "final", "obs/enc", "tbl loads", "tbl pairs", "tbl store", "table"]
    cols = [13, 5, 6, 7, 8, 9, 10, 11, 12, 0, 2, 3, 4, 1]
    print("cumulative us at the end of each phase:", names)
    for lo, hi in ((1, 4), (5, 8), (9, 12), (13, 16), (17, 24), (25, 32)):
        q = (nn >= lo) & (nn <= hi)
        if q.any():
            pm = ph[m][q].mean(axis=0) / 1e3
            print(f"nodes {lo}-{hi} (edges {(ne[m][q] & 0xFFFF).mean():.1f}):", " ".join(f"{pm[c]:.2f}" for c in cols), f"total {dur[m][q].mean() / 1e3:.2f}")
    o = np.lexsort((start, wid))
    same = wid[o][1:] == wid[o][:-1]
    gap = (start[o][1:] - (start[o][:-1] + dur[o][:-1]))[same]
    print("gap between items of a warp us: mean", gap.mean() / 1e3, "p50", np.percentile(gap, 50) / 1e3, "p99", np.percentile(gap, 99) / 1e3, "sum per warp", gap.sum() / nw / 1e3)
