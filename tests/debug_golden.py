"""TEST INFRASTRUCTURE (not collected by pytest) - debug aid: replay a golden case on the GPU next to the oracle and dump both states at the first divergence.
    python tests/debug_golden.py d12_reimage_random 1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccbs_b200 import lib as L                                  # noqa: E402
from ccbs_b200.batched_env import BatchedCyberBattleEnv         # noqa: E402
from oracle import gen_golden as gg, trace as tr                # noqa: E402
from oracle.cbs_oracle import OracleEnv                         # noqa: E402
from tests.gpu_harness import masks_to_u64                      # noqa: E402

name, gemm = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
case = gg.load_case(os.path.join(ROOT, "tests", "golden", name + ".npz"))
spec, cfg = case["spec"], case["cfg"]
B = 2
interest = None if case["interest"] is None else [case["interest"]]
env = BatchedCyberBattleEnv([spec], case["weights"], cfg, num_envs=B, auto_reset=True, decode_gemm=gemm, interest_nodes=interest,
                            decode_margin=float(sys.argv[3]) if len(sys.argv) > 3 else 0.0)
env.set_starter_queue(np.tile(case["starters"][None, :], (B, 1)))
o = OracleEnv(spec, case["weights"], cfg, interest_node=case["interest"])
vidx = tr.vuln_index(spec)
dd = case["defender_draws"]
if dd is not None:
    k = dd[0].shape[1]
    sn = torch.zeros((B, k), dtype=torch.int32, device=env.device)
    su = torch.zeros((B, k), dtype=torch.float32, device=env.device)
    env.set_defender_draws(sn, su)
ep = 0
o.reset(starter=int(case["starters"][0]))
env.reset()
env.sync()
actions = np.array(case["actions"], copy=True)
for t in range(len(actions)):
    if case["policy_rows"] is not None:
        actions[t] = (np.asarray(o.action_rows[int(case["policy_rows"][t])], np.float64) + actions[t].astype(np.float64)).astype(np.float32)
    a = torch.from_numpy(actions[t]).to(env.device).unsqueeze(0).repeat(B, 1).contiguous()
    u = torch.full((B,), float(case["uniforms"][t]), dtype=torch.float32, device=env.device)
    sel, dist = env.decode(a)
    env.sync()
    s, t_, vid, kind, d, row = o.find_closest_action_embedding(actions[t])
    want = (s, t_, vidx[vid], kind)
    got = tuple(int(x) for x in sel[0].cpu().numpy())
    if got != want:
        print(f"step {t} (episode {ep}, stepcount {o.stepcount}): oracle {want} d={d:.9f} row {row}; gpu {got} d={float(dist[0]):.9f}")
        print(" oracle owned_nodes", o.owned_nodes, "discovered", o.discovered_nodes, "reimaging", o.reimaging, "stale", o._stale)
        print(" oracle status", [int(n.status) for n in o.nodes], "installed", [int(n.agent_installed) for n in o.nodes])
        print(" oracle processed pairs with s ==", s, sorted(p for p in o.processed_pairs if p[0] == s))
        sc = env.scalars()
        nraw = int(sc[L.S_N_OWNED_RAW, 0])
        print(" gpu owned_order", env.owned_order()[0, :int(sc[L.S_N_OWNED, 0])].tolist(),
              "raw", env.owned_raw()[0, :nraw].tolist() if dd is not None else None,
              "disc", env.disc_order()[0, :int(sc[L.S_N_DISC, 0])].tolist(), "slots", int(sc[L.S_N_SLOTS, 0]),
              "flags", hex(int(sc[L.S_FLAGS, 0])), "encodes", int(sc[L.S_N_ENCODES, 0]), "oracle encodes", o.n_encodes)
        ps = env.read(L.F_PAIR_SLOT, np.uint8, (B, env.ncap, env.ncap))[0]
        print(" gpu pair_slot row s:", ps[s, :spec.num_nodes].tolist())
        zh = env.read(L.F_Z_HIST, np.float32, (B, env.slots, env.ncap, 64))[0]
        sl_ = int(ps[s, t_])
        orow = np.asarray(o.action_rows[row])
        print(" gpu slot", sl_, "z[s] vs oracle es max diff", np.abs(zh[sl_, s] - orow[:64]).max(), " z[t] vs et", np.abs(zh[sl_, t_] - orow[64:128]).max())
        for k_ in range(int(sc[L.S_N_SLOTS, 0])):
            print("   slot", k_, "diff s", np.abs(zh[k_, s] - orow[:64]).max(), "diff t", np.abs(zh[k_, t_] - orow[64:128]).max())
        print(" gpu masks", [hex(int(x)) for x in masks_to_u64(env.masks(), 0)[:, 0]])
        print(" ora masks", [hex(int(x)) for x in tr.masks_to_array(o.masks())[:, 0]])
        break
    draws = None if dd is None else (dd[0][t], dd[1][t])
    o.step(actions[t], case["uniforms"][t], defender_draws=draws)
    if dd is not None:
        sn.copy_(torch.from_numpy(np.tile(dd[0][t][None, :], (B, 1))))
        su.copy_(torch.from_numpy(np.tile(dd[1][t][None, :], (B, 1))))
    env.transition(sel, dist, u)
    env.sync()
    gm, om = masks_to_u64(env.masks(), 0), tr.masks_to_array(o.masks())
    obs = env.observe()
    env.sync()
    if not (o.done or o.truncated):
        og = np.concatenate([o.observation["graph_embeddings"], o.observation["discrete_features"]]).astype(np.float32)
        gg_ = obs[0].cpu().numpy()
        if not np.allclose(gg_, og, rtol=1e-5, atol=2e-5):
            print(f"step {t} (episode {ep}, stepcount {o.stepcount}): obs differ max {np.abs(gg_ - og).max():.3e}; sel {want} code {o.outcome}")
            print(" gpu tail", gg_[-2:], "oracle tail", og[-2:])
            print(" oracle owned", o.owned_nodes, "reimaging", o.reimaging, "graph nodes", o.graph_nodes,
                  "x status", {n: int(o.node_x[n][37]) for n in o.graph_nodes}, "live", [int(n.status) for n in o.nodes])
            sc = env.scalars()
            print(" gpu flags", hex(int(sc[L.S_FLAGS, 0])), "masks", [hex(int(x)) for x in gm[:, 0]])
            break
    if not np.array_equal(gm, om):
        print(f"step {t}: masks differ after the step: sel {want} code {o.outcome}")
        print(" gpu", [hex(int(x)) for x in gm[:, 0]])
        print(" ora", [hex(int(x)) for x in om[:, 0]])
        break
    if o.done or o.truncated:
        ep += 1
        o.reset(starter=int(case["starters"][ep]))
else:
    print("no divergence")
