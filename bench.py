#!/usr/bin/env python
"""bench.py — env-steps/s of the batched continuous env on B200, next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1] [--impl ours|reference]

One "step" = one pass of the hot path (decode -> transition -> observe, finished envs reset in place) over
every env of every rank.  Workload c2 (default) is BASELINE.json configs[2]: 65 536 envs of 32-node synthetic
scenarios sharded 8192 per GPU (weak scaling: N GPUs step N x 8192 envs); c1 is configs[1]: 4096 envs of
10-25-node scenarios on one GPU.  Actions are U(-4,4)^905 (the reference's action_space.sample()), pre-staged
in HBM as a ring of batches larger than L2.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(name="configs[2]: 65536 envs x 32-node synthetic scenarios, sharded 8192 envs/GPU (weak scaling)",
               envs_per_gpu=8192, nodes=(32, 32), scenarios=16),
    "c1": dict(name="configs[1]: 4096 envs x 10-25-node synthetic scenarios on 1 GPU", envs_per_gpu=4096, nodes=(10, 25),
               scenarios=16),
    "cd": dict(name="default-like deployment set (docs/ch2_env_stats.md): 20 scenarios x 100 nodes, ~33 vulnerabilities/node, "
                    "pool of 3000 CVEs, 4096 envs", envs_per_gpu=4096, nodes=(100, 100), scenarios=20, pool=3000,
               services_range=(1, 4), vulns_per_service_range=(8, 20)),
    "c4": dict(name="configs[3]: mixed-topology batch, 10-100-node scenarios (padded to 100 nodes), 768-d embeddings, pool 600",
               envs_per_gpu=4096, nodes=(10, 100), scenarios=20, pool=600),
}
POOL_SEED, GAE_SEED = 1234, 0


def scenario_params(wl):
    """(graph seed, node count, generator kwargs) of every scenario of a workload"""
    rng = np.random.default_rng(2024)
    gkw = {k2: wl[k2] for k2 in ("services_range", "vulns_per_service_range") if k2 in wl}
    return [(100 + k, int(rng.integers(wl["nodes"][0], wl["nodes"][1] + 1)), gkw) for k in range(wl["scenarios"])]


def build_specs(wl):
    import ccbs_b200 as cb
    pool = cb.synthetic_vuln_pool(POOL_SEED, wl.get("pool", 200))
    return [cb.synthetic_spec(seed, n, pool=pool, **gkw) for seed, n, gkw in scenario_params(wl)]


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) — the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wl_key, worker, seconds = args
    import torch
    torch.set_num_threads(1)
    import ccbs_b200 as cb
    from ccbs_b200.gae import GaeWeights
    from oracle.cbs_oracle import OracleEnv
    wl = WORKLOADS[wl_key]
    specs = build_specs(wl)
    spec = specs[worker % len(specs)]
    env = OracleEnv(spec, GaeWeights.random(GAE_SEED), cb.EnvConfig())
    rng = np.random.default_rng(1000 + worker)
    env.reset(rng=rng)
    steps = resets = 0
    t_reset = 0.0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = rng.uniform(-4, 4, size=905).astype(np.float32)
        _, _, done, _ = env.step(a, rng.random())
        steps += 1
        if done:
            t1 = time.perf_counter()
            env.reset(rng=rng)
            t_reset += time.perf_counter() - t1
            resets += 1
    return steps, resets, time.perf_counter() - t0, t_reset


def _ref_worker(args):
    """One process stepping the UNMODIFIED reference env (oracle/_ref bytecode, or /root/reference where it is mounted): the
    reference's own scenario generator on the workload's synthetic input graph, its own RandomSwitchEnv over one
    CyberBattleCompressedEnv, its own action_space.sample() actions and global `random` draws, auto-reset like DummyVecEnv."""
    wl_key, worker, seconds = args
    import torch
    torch.set_num_threads(1)
    import ccbs_b200 as cb
    from ccbs_b200.gae import GaeWeights
    from oracle import ref_bridge as rb
    wl = WORKLOADS[wl_key]
    params = scenario_params(wl)
    seed, n, gkw = params[worker % len(params)]
    pool = cb.synthetic_vuln_pool(POOL_SEED, wl.get("pool", 200))
    graph = cb.synthetic_input_graph(seed, n, pool=pool, **gkw)
    model = rb.reference_model_from_input_graph(graph, seed=seed)
    env = rb.make_unpatched_env(model, GaeWeights.random(GAE_SEED), cb.EnvConfig(), seed=1000 + worker)
    env.reset()
    steps = resets = 0
    t_reset = 0.0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, trunc, _ = env.step(env.action_space.sample())
        steps += 1
        if done or trunc:
            t1 = time.perf_counter()
            env.reset()
            t_reset += time.perf_counter() - t1
            resets += 1
    return steps, resets, time.perf_counter() - t0, t_reset


def reference_kind():
    """'reference' when the unmodified reference can be imported here (oracle/_ref, built by oracle/build_ref.py), else 'port'"""
    from oracle import ref_bridge as rb
    return "reference" if rb.reference_available() else "port"


def cpu_baseline(wl_key, seconds, procs, kind="port"):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker if kind == "reference" else _cpu_worker, [(wl_key, w, seconds) for w in range(procs)])
    steps = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return dict(value=steps / wall, steps=steps, wall_s=wall, resets=sum(r[1] for r in res),
                reset_frac=sum(r[3] for r in res) / sum(r[2] for r in res))


def cpu_sample_text(kind, procs, seconds, res):
    what = ("the unmodified reference RandomSwitchEnv(CyberBattleCompressedEnv) (oracle/_ref), its own generator on the workload's "
            "input graphs, action_space.sample() actions" if kind == "reference" else "one oracle-port env (OracleEnv) each on the same scenario set, random actions")
    return (f"{procs} processes x {seconds:.0f}s, {what}, auto-reset ({res['steps']} steps, {res['resets']} resets, "
            f"{100 * res['reset_frac']:.0f}% of time in reset)")


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ready = threading.Event()
        self.armed = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason")}
            self.ready.set()
            while not self.stop_flag:
                if not self.armed:
                    time.sleep(0.001)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit:
                        n = name.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                        if n not in ("None", "All", "GpuIdle"):
                            self.reasons.add(n)
                time.sleep(self.period)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_error:{type(e).__name__}")
            self.ready.set()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the last committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/summarize_ncu.py); None if that kernel was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel)


def transition_roofline(specs, weights, device_index, n_envs, n_steps=8, seed=11, blocked=None):
    """Large-batch legs for the transition kernel (at 8192 envs a launch moves 1.8 MB and is latency bound; the
    HBM roofline is only meaningful at >= 1e6 env-steps per launch).  The same scenario set with `n_envs` envs on one GPU.
    (1) per-step launch: the split C-ABI calls decode -> transition -> observe from a fresh reset, only `cbs_transition` timed
    (CUDA events on the launching stream).  (2) K-step persistent launch: the decoded actions, distances and uniforms of those
    very steps are replayed from the same reset state by ONE `cbs_transition_ksteps` launch (records in registers across the
    steps, written back once).  Snapshot capacity is cut to 4 slots so that the state fits HBM; envs that would need more stop
    growing their tables (flagged, harmless for this measurement).  Returns a dict."""
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import constants as C
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    dev = torch.device("cuda", device_index)
    sc_of_env = None
    if blocked if blocked is not None else bool(int(os.environ.get("CBS_BLOCKED", "0"))):
        sc_of_env = ((np.arange(n_envs, dtype=np.int64) * len(specs)) // n_envs).astype(np.int32)
    env = BatchedCyberBattleEnv(specs, weights, cb.EnvConfig(), num_envs=n_envs, device=device_index, seed=seed,
                                auto_reset=True, max_slots=4, max_edges=8, scenario_of_env=sc_of_env)
    # every episode of env b starts from the first feasible starter of its scenario: a reset reproduces the same state
    g = C.GOALS["control"]
    t = env.tables
    first = np.asarray(t.feasible_starters[g])[np.asarray(t.sc_feasible_off[g])[:-1]]
    env.set_starter_queue(first[env.scenario_of_env][:, None].astype(np.int32))
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    actions = torch.empty(n_envs, C.ACTION_DIM, device=dev)
    K = int(n_steps)
    sel_all = torch.empty(K, n_envs, 4, dtype=torch.int32, device=dev)
    dist_all = torch.empty(K, n_envs, dtype=torch.float64, device=dev)
    uni_all = torch.rand(K, n_envs, device=dev, generator=gen)
    rew_all = torch.empty(K, n_envs, dtype=torch.float32, device=dev)
    env.reset()
    torch.cuda.synchronize(dev)
    times, ok_frac = [], []
    alive = torch.ones(n_envs, dtype=torch.bool, device=dev)     # envs whose first episode is still running
    alive_frac = []
    for k in range(K):            # step 0 is the warm-up launch
        actions.uniform_(-4.0, 4.0, generator=gen)
        sel, dist = env.decode(actions)
        sel_all[k].copy_(sel)
        dist_all[k].copy_(dist)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reward, done, _, outcome = env.transition(sel, dist, uni_all[k])
        e1.record()
        rew_all[k].copy_(reward)
        env.observe()
        torch.cuda.synchronize(dev)
        alive_frac.append(float(alive.float().mean().item()))
        alive &= done == 0
        if k:
            times.append(e0.elapsed_time(e1))
            ok_frac.append(float((outcome < 16).float().mean().item()))
    state_gb = env.state_bytes / 1e9
    ms = float(np.mean(times))
    peak, peak_src = measured_peaks()
    alg = 220.0 * n_envs                    # ~220 B of per-env state and I/O per transition
    out = {"kernel": "transition (split call cbs_transition, large batch)", "envs_per_launch": int(n_envs), "launches_timed": len(times),
           "ms_per_launch": ms, "ms_all": times, "bytes_per_env_step": 220, "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
           "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
           "traffic": ncu_traffic("transition_large"), "successful_outcomes": float(np.mean(ok_frac)),
           "state_gb": state_gb, "env_steps_per_s": n_envs / (ms * 1e-3)}
    # ---- (2) the same K steps as one persistent launch ----
    try:
        kms = []
        same = None
        for rep in range(6):
            env.reset()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rk, dk = env.transition_ksteps(sel_all, dist_all, uni_all)
            e1.record()
            torch.cuda.synchronize(dev)
            if rep:
                kms.append(e0.elapsed_time(e1))
            else:   # while an env's first episode runs, the persistent launch must reproduce the per-step launches bit for bit
                running = torch.ones(n_envs, dtype=torch.bool, device=dev)
                same = True
                for k in range(K):
                    same = same and bool(torch.equal(rk[k][running], rew_all[k][running]))
                    running &= dk[k] == 0
        kms_mean = float(np.median(kms))      # (the first launches after the resets of 1M envs still see allocator / page effects)
        # physical bytes the launch needs per env: records once (hot sector 32 in + 32 out, list lengths 16, masks 64 in + changed
        # words out) + per step sel 16 + dist 8 + uniform 4 + reward 4 + done 1
        phys = 150.0 + 33.0 * K
        out["ksteps"] = {"kernel": "transition_ksteps (K steps per launch, records in registers)", "k_steps": K, "envs_per_launch": int(n_envs),
                         "ms_per_launch": kms_mean, "ms_all": kms, "env_steps_per_s": n_envs * K / (kms_mean * 1e-3),
                         "bytes_per_env_step": 220, "achieved": alg * K / (kms_mean * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg * K / (kms_mean * 1e-3) / 1e9 / peak,
                         "physical_bytes_per_env_step_model": phys / K, "physical_gbs_model": phys * n_envs / (kms_mean * 1e-3) / 1e9,
                         "matches_per_step_launches": same, "envs_running_at_step": alive_frac,
                         "traffic": ncu_traffic("transition_ksteps")}
    except Exception as exc:  # noqa: BLE001
        out["ksteps"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    env.close()
    del actions, sel_all, dist_all, uni_all, rew_all
    torch.cuda.empty_cache()
    return out


def side_workload(key, device_index, steps=60, **cfg_kw):
    """Device-resident env-steps/s of another workload (same measurement as the headline, plain launches, short): BASELINE
    configs[1] / configs[3] next to the headline configs[2], and the headline workload with reference options switched on."""
    import torch
    import ccbs_b200 as cb
    from ccbs_b200 import constants as C
    wl = WORKLOADS[key]
    specs = build_specs(wl)
    B = wl["envs_per_gpu"]
    dev = torch.device("cuda", device_index)
    env = cb.BatchedCyberBattleEnv(specs, cb.GaeWeights.random(GAE_SEED), cb.EnvConfig(**cfg_kw), num_envs=B, device=device_index, seed=7)
    R = max(2, int(np.ceil(160e6 / (B * C.ACTION_DIM * 4))))
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    ring = torch.rand(R, B, C.ACTION_DIM, device=dev, generator=gen) * 8.0 - 4.0
    env.reset()
    stagger = torch.arange(B, device=dev) % 32
    for i in range(64):
        if i < 32:
            env.reset((stagger == i).to(torch.uint8))
        env.step(ring[i % R], None, want_info=False)
    env.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.step(ring[i % R], None, want_info=False)
    e1.record()
    env.sync()
    ms = e0.elapsed_time(e1)
    out = {"workload": wl["name"], "envs": B, "options": cfg_kw, "steps": steps, "ms_per_step": ms / steps,
           "env_steps_per_s": B * steps / (ms * 1e-3), "state_gb": env.state_bytes / 1e9, "decode_margin_edge": env.margin_edge_count()}
    env.close()
    del ring
    torch.cuda.empty_cache()
    return out


def configs4_leg(device_index, n_envs=4096, steps=40):
    """BASELINE configs[4] (SB3 PPO with the GPU env in the loop).  stable_baselines3 is not in this image, so two things are timed:
    (1) the SB3 boundary itself — ``CyberBattleVecEnv.step(numpy actions) -> (dict obs, rewards, dones, infos)``, the call
    SB3's collect_rollouts makes (agents/train_agent.py:113) — with full and with lazy info dicts; (2) the in-repo PPO loop of
    examples/train_ppo_b200.py (rollout + update on the same GPU): env-steps/s and wall-clock per update.  The reference's
    figure for the same boundary is the CPU arm's env-steps/s (one DummyVecEnv env per process)."""
    import importlib.util
    import ccbs_b200 as cb
    from ccbs_b200 import constants as C
    from ccbs_b200.vec_env import CyberBattleVecEnv
    wl = WORKLOADS["c1"]
    specs = build_specs(wl)
    weights = cb.GaeWeights.random(GAE_SEED)
    out = {"envs": n_envs, "workload": wl["name"], "steps": steps}
    rng = np.random.default_rng(0)
    a = rng.uniform(-4, 4, size=(n_envs, C.ACTION_DIM)).astype(np.float32)
    for lazy in (False, True):
        venv = CyberBattleVecEnv(cb.BatchedCyberBattleEnv(specs, weights, cb.EnvConfig(), num_envs=n_envs, device=device_index, seed=3),
                                 lazy_infos=lazy)
        venv.reset()
        for _ in range(5):
            venv.step(a)
        t0 = time.perf_counter()
        n_done = 0
        for _ in range(steps):
            _, _, dones, _ = venv.step(a)
            n_done += int(dones.sum())
        dt = time.perf_counter() - t0
        out["vecenv_step_lazy_infos" if lazy else "vecenv_step_full_infos"] = {
            "env_steps_per_s": n_envs * steps / dt, "ms_per_step": 1e3 * dt / steps, "episode_ends": n_done}
        venv.close()
    spec = importlib.util.spec_from_file_location("train_ppo_b200", os.path.join(ROOT, "examples", "train_ppo_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    log = mod.main(["--envs", str(n_envs), "--updates", "4", "--n-steps", "32"], quiet=True)
    last = log[-1]
    out["ppo_in_repo"] = {"rollout_env_steps_per_s": last["rollout_env_steps_per_s"], "train_env_steps_per_s": last["train_env_steps_per_s"],
                          "seconds_per_update": last["seconds_per_update"], "env_steps_per_update": 32 * n_envs,
                          "policy": "MLP [256,128,64] actor + critic, 4 epochs x 16384-sample minibatches, same GPU"}
    return out


def run_reference(args, wl_key, rank, emit):
    """--impl reference: the CPU restatement of the same path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    wl = WORKLOADS[wl_key]
    kind = reference_kind()
    # every window spawns its processes afresh (the real reference spends ~20 s importing and generating its scenario before
    # it steps), so the K "steps" of this arm are a few longer windows rather than K short ones
    n_win = max(1, min(args.steps, 3 if kind == "reference" else args.steps))
    per_step = max(1.0, min(8.0 if kind == "reference" else 6.0, 60.0 / n_win))
    vals = []
    t0 = time.time()
    for _ in range(n_win):
        vals.append(cpu_baseline(wl_key, per_step, procs, kind))
        if time.time() - t0 > 150:
            break
    steps = sum(v["steps"] for v in vals)
    wall = sum(v["wall_s"] for v in vals)
    v = steps / wall
    sample = f"{len(vals)} windows: " + cpu_sample_text(kind, procs, per_step, dict(steps=steps, resets=sum(x["resets"] for x in vals),
                                                                                   reset_frac=float(np.mean([x["reset_frac"] for x in vals]))))
    emit({
        "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 (python)", "data": "synthetic",
        "config": {"workload": wl["name"]},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to its GPU's local CPU cores")
    ap.add_argument("--ring", type=int, default=0, help="minimum number of action batches in the device-resident ring")
    ap.add_argument("--no-graph", action="store_true", help="time plain per-step launches instead of CUDA-graph replays")
    ap.add_argument("--no-prestaged", action="store_true", help="do not declare the action ring pre-staged (cbs_set_actions_prestaged): the "
                    "contraction then waits for the previous step's observe kernel before reading its actions, as it must behind a policy")
    ap.add_argument("--presteps", type=int, default=64, help="untimed steps before the warm-up: episodes reach their steady-state mix")
    ap.add_argument("--subset", type=int, default=0, help="side measurement: sample_subset_samples=K (the reference's training default is 100)")
    ap.add_argument("--no-vecenv", action="store_true", help="skip the configs[4] leg (CyberBattleVecEnv.step and the PPO loop)")
    ap.add_argument("--decode-gemm", type=int, default=0)
    ap.add_argument("--defender", action="store_true",
                    help="side measurement: the same workload with the re-imaging static defender of train_config.yaml:39-44 "
                         "(not the headline configuration)")
    ap.add_argument("--host-shards", type=int, default=4, help="handles the batch is cut into for the pipelined end-to-end leg")
    ap.add_argument("--host-shard-weights", type=str, default="", help="comma-separated relative slice sizes of the pipelined end-to-end leg "
                    "(overrides --host-shards; a small last slice shortens the tail after the action copy)")
    ap.add_argument("--transition-envs", type=int, default=1 << 20,
                    help="envs of the large-batch transition-kernel roofline leg (rank 0, N=1 only; 0 = skip)")
    ap.add_argument("--action-pitch", type=int, default=905,
                    help="row pitch (floats) of the device-resident action batches; 905 = dense like the reference's "
                         "action array, 908 lets TMA read them in place (no repack kernel)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 meanwhile (e.g. NCCL's version banner)
    # is diverted to stderr until the result is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    # Every action batch this script steps with is generated before the timed window (a ring of device tensors, or pinned host
    # buffers copied by the library itself): declare that to the library (include/cbsim.h: cbs_set_actions_prestaged), so that
    # step t+1's contraction may read its actions while step t's observe kernel drains.  --no-prestaged measures without it.
    if not args.no_prestaged:
        os.environ.setdefault("CBS_ACTIONS_PRESTAGED", "1")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.envs_per_gpu:
        wl["envs_per_gpu"] = args.envs_per_gpu
    if args.impl == "reference":
        run_reference(args, args.workload, rank, emit)
        return

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        kind = reference_kind()
        res = cpu_baseline(args.workload, args.cpu_seconds, procs, kind)
        cpu = {"value": res["value"], "unit": "env-steps/s", "cores": procs, "kind": kind,
               "sample": cpu_sample_text(kind, procs, args.cpu_seconds, res)}
        if kind == "reference":     # the oracle port beside it: the restatement is ~25x faster per core than the reference it restates
            res_p = cpu_baseline(args.workload, min(args.cpu_seconds, 6.0), procs, "port")
            cpu["port_value"] = res_p["value"]
            cpu["port_sample"] = cpu_sample_text("port", procs, min(args.cpu_seconds, 6.0), res_p)

    import torch
    import torch.distributed as dist
    import ccbs_b200 as cb
    from ccbs_b200 import constants as C
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from ccbs_b200.gae import GaeWeights

    # host side of the end-to-end leg: run this rank on the cores next to its GPU, so that the pinned host buffers (first touch)
    # and the launching thread sit on the GPU's own NUMA node and PCIe root complex
    affinity = None
    if not args.no_affinity:
        try:
            import pynvml as nv
            nv.nvmlInit()
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local_rank))
            affinity = len(os.sched_getaffinity(0))
        except Exception as exc:  # noqa: BLE001
            affinity = f"unavailable ({type(exc).__name__})"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = wl["envs_per_gpu"]
    specs = build_specs(wl)
    weights = GaeWeights.random(GAE_SEED)
    cfg = cb.EnvConfig(static_defender_agent="reimage") if args.defender else cb.EnvConfig(sample_subset_samples=args.subset)
    env = BatchedCyberBattleEnv(specs, weights, cfg, num_envs=B, device=local_rank, seed=7, global_env_offset=rank * B,
                                auto_reset=True, decode_gemm=args.decode_gemm)
    # action ring: R batches of [B, 905] float32, together larger than the 126 MB L2 (step i reads batch i % R).  The timed steps
    # are replayed from a CUDA graph of G consecutive steps when --steps has an even divisor G in [R, 32] (the driver's 20
    # steps: one graph of 20), so that one replay walks the whole ring more than once.  (A longer ring only evicts more of the
    # env state from L2 between steps — measured 63 / 56 / 51 M env-steps/s with 6 / 10 / 20 batches; a trainer's actions are
    # the freshly written output of its policy network, not a ring, so the shortest ring that exceeds L2 is the one used.)
    R = max(2, int(np.ceil(160e6 / (B * C.ACTION_DIM * 4))), args.ring)
    G = 0
    if not args.no_graph:
        for cand in range(min(32, args.steps), R - 1, -1):
            if cand % 2 == 0 and args.steps % cand == 0:
                G = cand
                break
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    ring = (torch.rand(R, B, args.action_pitch, device=dev, generator=gen) * 8.0 - 4.0).contiguous()[:, :, :C.ACTION_DIM]
    env.reset()
    env.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from here on: the pre-steps below are the same load as the (possibly very short) timed window
    sampler = ClockSampler(local_rank, period=0.002)
    sampler.start()
    sampler.ready.wait(10)
    sampler.armed = True
    # bring the episodes to their steady-state mix before anything is timed (a window that starts right after a cold reset
    # of all envs holds no episode end and no in-place reset), then the W warm-up steps of the contract
    # (all envs start together and the proportional cut-off ends most episodes after the same ~32 steps: without the staggered
    # resets below the batch stays in phase and a 20-step window either contains the whole burst of episode ends or none of it)
    stagger = torch.arange(B, device=dev) % 32
    for i in range(args.presteps):
        if i < 32:
            env.reset((stagger == i).to(torch.uint8))
        env.step(ring[i % R], None, want_info=False)
    env.sync()
    env.reset_stat_accum()
    for i in range(args.warmup):
        env.step(ring[i % R], None, want_info=False)
    env.sync()
    graph = None
    if G:
        try:      # G steps of the same three launches each, captured once on a side stream, replayed steps / G times
            gstream = torch.cuda.Stream(device=dev)
            gstream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(gstream):
                for i in range(2):                       # untimed: nothing lazy (function attributes, tensor maps) is left for the capture
                    env.step(ring[i % R], None, want_info=False)
            torch.cuda.current_stream().wait_stream(gstream)
            torch.cuda.synchronize()
            l_before = env.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=gstream):
                for i in range(G):
                    env.step(ring[i % R], None, want_info=False)
            launches_per_replay = env.launch_count - l_before
            graph.replay()                               # one untimed replay
            env.sync()
        except Exception as exc:  # noqa: BLE001  (a failed capture must not cost the measurement: plain launches instead)
            print(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {exc}); timing plain launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    # ---- timed region: device-resident inputs ----
    barrier()
    l0 = env.launch_count
    launches_per_replay = 0 if graph is None else launches_per_replay
    side = torch.cuda.Stream(device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        for _ in range(args.steps // G):
            graph.replay()
    else:
        for i in range(args.steps):
            env.step(ring[(args.warmup + i) % R], None, want_info=False)
    e1.record()
    # the only collective: episode statistics, once per logging interval — issued on a side stream behind the window's last
    # step, where a trainer overlaps it with the next interval's steps
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        acc_all = env.stat_accum_tensor().clone()
        if world > 1:
            dist.all_reduce(acc_all)
    torch.cuda.current_stream().wait_stream(side)
    barrier()
    launches = (args.steps // G) * launches_per_replay if graph is not None else env.launch_count - l0
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    env.sync()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total_envs = B * world
    value = total_envs * args.steps / (ms * 1e-3)

    # ---- end-to-end through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region) ----
    # (a) one handle, cbs_step_host: copy in, step, copy out, wait.  (b) the same batch held by 4 handles stepped as a
    # pipeline (ShardedHostEnv: slice k+1's copy-in runs under slice k's kernels and copy-out) — the reported e2e.
    h_ring = [torch.empty(B, C.ACTION_DIM, dtype=torch.float32).pin_memory() for _ in range(2)]
    for k in range(2):
        h_ring[k].copy_(ring[k].cpu().contiguous())
    h_obs = torch.empty(B, C.OBS_DIM + 2, dtype=torch.float32).pin_memory()
    h_rew = torch.empty(B, dtype=torch.float32).pin_memory()
    h_done = torch.empty(B, dtype=torch.uint8).pin_memory()
    e2e_steps = max(10, args.steps // 2)

    def time_host_steps(stepper):
        for i in range(3):
            stepper(h_ring[i % 2].numpy(), None, h_obs.numpy(), h_rew.numpy(), h_done.numpy())
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            stepper(h_ring[i % 2].numpy(), None, h_obs.numpy(), h_rew.numpy(), h_done.numpy())
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return total_envs * e2e_steps / float(tt.item())

    # what the link alone allows: the same pinned action buffer copied to the device, nothing else (CUDA events)
    d_probe = torch.empty(B, C.ACTION_DIM, dtype=torch.float32, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        d_probe.copy_(h_ring[0], non_blocking=True)
    torch.cuda.synchronize()
    ev0.record()
    for k in range(20):
        d_probe.copy_(h_ring[k % 2], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_ms = ev0.elapsed_time(ev1) / 20
    del d_probe
    # the same probe with ALL ranks copying at once: what the host's memory system / PCIe root complexes deliver per GPU when
    # every rank streams its action batch (the end-to-end leg's situation at N > 1)
    h2d_all = [h2d_ms]
    if world > 1:
        d_probe = torch.empty(B, C.ACTION_DIM, dtype=torch.float32, device=dev)
        barrier()
        ev0.record()
        for k in range(20):
            d_probe.copy_(h_ring[k % 2], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([ev0.elapsed_time(ev1) / 20], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        h2d_all = [float(x.item()) for x in gathered]
        del d_probe
    e2e_single = time_host_steps(env.step_host)
    from ccbs_b200.host_pipeline import ShardedHostEnv
    e2e_api = f"ShardedHostEnv.step_host: {args.host_shards} handles x cbs_step_host_async, pinned host buffers"
    try:
        sw = [float(x) for x in args.host_shard_weights.split(",")] if args.host_shard_weights else None
        if sw:
            e2e_api = f"ShardedHostEnv.step_host: {len(sw)} handles (slice weights {args.host_shard_weights}) x cbs_step_host_async, pinned host buffers"
        penv = ShardedHostEnv(specs, weights, cfg, num_envs=B, shards=args.host_shards, shard_weights=sw, device=local_rank, seed=7,
                              global_env_offset=rank * B, auto_reset=True, decode_gemm=args.decode_gemm)
        penv.reset()
        for i in range(64):      # bring the episodes to their steady-state mix before timing
            for (lo, hi), e in zip(penv.bounds, penv.envs):
                e.step(ring[i % R][lo:hi], None, want_info=False)
        penv.sync()
        e2e_value = time_host_steps(penv.step_host)
        e2e_launches_per_step = 3 * len(penv.envs)
        penv.close()
    except Exception as exc:  # noqa: BLE001  (fall back to the single-handle measurement rather than lose the line)
        e2e_value, e2e_launches_per_step = e2e_single, 3
        e2e_api = f"cbs_step_host, one handle (pipelined leg failed: {type(exc).__name__})"
    h2d = B * C.ACTION_DIM * 4
    d2h = B * ((C.OBS_DIM + 2) * 4 + 4 + 1)

    # ---- per-kernel durations: CUDA events between the launches, on the launching stream (cbs_profile_step) ----
    n_prof = min(50, args.steps)
    kern = {}
    for i in range(3):   # untimed: the split-call kernel variants are loaded lazily on first use
        env.profile_step(ring[i % R])
    for i in range(n_prof):
        ms_k = env.profile_step(ring[i % R])
        for k, v in ms_k.items():
            kern[k] = kern.get(k, 0.0) + v / n_prof
    env.sync()

    if rank == 0:
        peak, peak_src = measured_peaks()
        # algorithmic bytes per env-step: decode = action read 3620 B + VT row write+read
        # 2*4*Ug + candidate scan; transition = 220 B; observe = obs write 776 B + (on re-encode) ~10 KB
        from ccbs_b200 import lib as _L
        Ug = env.tables.vemb32.shape[0]
        sc = env.scalars()
        pairs = float(np.mean(sc[_L.S_N_OWNED] * sc[_L.S_N_DISC]))   # n_owned * n_disc  (upper bound of table pairs)
        rows = pairs * 24.0                                         # ~24 candidate rows per pair (tools/workload_stats.py)
        alg = {
            # A_v read + Vemb + VT write (the dense action rows are staged by cp.async, no repack pass)
            "decode_gemm": B * 768 * 4 + Ug * 768 * 4 + B * Ug * 4,
            # action read + VT row read + per pair two half-precision snapshot rows + norms, per row 16 B of table,
            # plus the transition's 220 B of per-env state
            "decode_select_transition": B * (C.ACTION_DIM * 4 + 4 * Ug + pairs * (2 * 128 + 9) + rows * 16 + 28 + 220.0),
        }
        dom = max((k for k in kern if k in alg), key=kern.get)
        # (observe has no algorithmic-bytes model worth the name — its items are latency chains over L2-resident tables — so
        # its figure is the DRAM traffic ncu measured for one launch, profiles/traffic.json, over the live duration)
        gbs = {k: alg[k] / (kern[k] * 1e-3) / 1e9 for k in kern if k in alg}
        if ncu_traffic("observe"):
            gbs["observe_dram_ncu"] = ncu_traffic("observe") / (kern["observe"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": alg[dom] / (kern[dom] * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "peak_source": peak_src, "traffic": ncu_traffic(dom),
                "kernels_ms": kern, "kernels_gbs": gbs}
        roof["frac"] = roof["achieved"] / peak
        acc = dict(zip(_L.ACCUM_NAMES, acc_all.cpu().tolist()))   # summed over all ranks, timed-region episodes + warm-up
        out = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bitmasks + f32 GAE + f64 decode re-score", "data": "synthetic",
            "config": {"workload": wl["name"] + (" + static defender (reimage 0.05/3/3)" if args.defender else "") +
                       (f" + sample_subset_samples={args.subset}" if args.subset else ""), "envs_per_gpu": B, "scenarios": wl["scenarios"], "global_vulns": int(Ug),
                       "action_pitch_floats": args.action_pitch,
                       "actions": f"ring of {R} x [{B},905] f32 batches = {R * B * 905 * 4 / 1e6:.0f} MB (> 126 MB L2), no L2 flush",
                       "launch": (f"CUDA graph of {G} steps replayed {args.steps // G}x" if graph is not None else "plain launches"),
                       "programmatic_dependent_launch": os.environ.get("CBS_NO_PDL") is None,
                       "actions_prestaged": os.environ.get("CBS_ACTIONS_PRESTAGED") is not None,
                       "presteps": args.presteps, "stats_allreduce": "side stream, behind the window's last step",
                       "sample_subset_samples": args.subset,
                       "decode_gemm": "tcgen05-tf32" if env.tensor_core_decode else "simt-f32",
                       "state_gb": env.state_bytes / 1e9},
            "clocks": sampler.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": e2e_api,
                    "gpu_launches_per_step": e2e_launches_per_step, "single_handle_value": e2e_single,
                    # the bare pinned copy of one rank's action batch, timed alone: what the link delivered on this box
                    "h2d_copy_ms": h2d_ms, "h2d_gbs": h2d / (h2d_ms * 1e-3) / 1e9, "cpu_affinity_cores": affinity,
                    # per rank, all ranks copying concurrently: the ceiling of the end-to-end leg at N GPUs is
                    # total_envs / max(copy time) — host memory / PCIe contention, not the kernels
                    "h2d_gbs_per_rank_concurrent": [round(h2d / (m * 1e-3) / 1e9, 2) for m in h2d_all],
                    "copy_bound_value": total_envs / (max(h2d_all) * 1e-3)},
            "roofline": roof, "cpu_baseline": cpu,
            "episodes": {k: acc[k] for k in ("episodes", "return_sum", "length_sum", "wins", "lost", "cutoff")},
            # decodes of the whole run whose float64 winner sat in the outer half of the float32 re-score margin (0 = the margin
            # was never under pressure)
            "decode_margin_edge": env.margin_edge_count(),
        }
        env.close()
        if world == 1 and not args.no_vecenv:
            try:
                out["configs4"] = configs4_leg(local_rank)
            except Exception as exc:  # noqa: BLE001
                out["configs4"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            # the other BASELINE configs and the reference's training options, measured the same way (short side legs)
            side = {}
            for name, key, kw in (("configs1", "c1", {}), ("configs3_mixed_10_100", "c4", {}),
                                  ("configs2_subset100", "c2", dict(sample_subset_samples=100)),
                                  ("configs2_reimage_defender", "c2", dict(static_defender_agent="reimage")),
                                  ("configs2_events_defender", "c2", dict(static_defender_agent="events"))):
                try:
                    side[name] = side_workload(key, local_rank, **kw)
                except Exception as exc:  # noqa: BLE001
                    side[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            out["other_workloads"] = side
        if world == 1 and args.transition_envs > 0:
            try:    # a side measurement must never cost the headline line (e.g. 60 GB of state not available on this box)
                out["transition_roofline"] = transition_roofline(specs, weights, local_rank, args.transition_envs)
            except Exception as exc:  # noqa: BLE001
                out["transition_roofline"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        emit(out)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
