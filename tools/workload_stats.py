import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ccbs_b200 as cb
from ccbs_b200 import lib as L, constants as C
from ccbs_b200.batched_env import BatchedCyberBattleEnv
from ccbs_b200.gae import GaeWeights
wl = bench.WORKLOADS['c2']; specs = bench.build_specs(wl)
B=2048
env = BatchedCyberBattleEnv(specs, GaeWeights.random(0), cb.EnvConfig(), num_envs=B, seed=7)
env.reset()
g = torch.Generator(device='cuda'); g.manual_seed(1)
for i in range(150):
    a = torch.rand(B, 905, device='cuda', generator=g)*8-4
    env.step(a, None, want_info=False)
env.sync()
sc = env.scalars(); ps = env.read(L.F_PAIR_SLOT, np.uint8, (B, env.ncap*env.ncap)).reshape(B, env.ncap, env.ncap)
T = env.tables
pairs = (ps != 255).sum(axis=(1,2))
rows = np.zeros(B)
for b in range(B):
    scn = sc[L.S_SCENARIO, b]; off = T.sc_node_off[scn]
    ss, tt = np.nonzero(ps[b] != 255)
    for s,t in zip(ss,tt):
        g_ = off + t
        r0 = T.nd_row_off[2*g_] if s==t else T.nd_row_off[2*g_+1]
        rows[b] += T.nd_row_off[2*g_+2]-r0
for name, v in [('n_disc', sc[L.S_N_DISC]), ('n_owned', sc[L.S_N_OWNED]), ('slots', sc[L.S_N_SLOTS]), ('edges', sc[L.S_N_EDGES]), ('pairs', pairs), ('rows', rows), ('stepcount', sc[L.S_STEPCOUNT])]:
    print(f"{name:10s} mean {np.mean(v):8.1f} p50 {np.percentile(v,50):7.1f} p90 {np.percentile(v,90):7.1f} max {np.max(v):7.1f}")
print('caps', env.ncap, env.slots, env.ecap, 'state GB', env.state_bytes/1e9)
