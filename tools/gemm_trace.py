"""Per-CTA timeline of the decode contraction (debug aid).  Needs k_decode_tc.cu compiled with -DCBS_GEMM_TRACE
(nvcc ... -DCBS_GEMM_TRACE -c k_decode_tc.cu, relink libcbsim.so); the stock library leaves the rows zero."""
import os, sys, ctypes as ct
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ccbs_b200 as cb
from ccbs_b200.batched_env import BatchedCyberBattleEnv
B = 8192
env = BatchedCyberBattleEnv(bench.build_specs(bench.WORKLOADS["c2"]), cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=7)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = [torch.rand(B, 905, device="cuda", generator=g) * 8 - 4 for _ in range(6)]
for i in range(40):
    env.step(acts[i % 6], None, want_info=False)
trace = torch.zeros(B * 6 + 128 * 8 * 6, dtype=torch.int64, device="cuda")
env.lib.cbs_debug_select_trace(env._h, ct.c_void_p(trace.data_ptr()))
env.step(acts[0], None, want_info=False); env.sync()
env.lib.cbs_debug_select_trace(env._h, None)
t = trace.cpu().numpy()[B * 6:].reshape(128, 8, 6)
t0 = t[:, :, 0][t[:, :, 0] > 0].min()
rel = np.where(t > 0, t - t0, 0) / 1e3
names = ["start", "tmem ready", "role loop done", "accumulator ready", "epilogue done", "after final barrier"]
for w, role in ((0, "warp 0 (TMA B)"), (1, "warp 1 (MMA)"), (2, "warp 2"), (4, "warp 4 (A producer)")):
    print(role, " ".join(f"{names[k]} {rel[:, w, k].mean():.2f} (max {rel[:, w, k].max():.2f})" for k in range(6)))
