"""TEST INFRASTRUCTURE (build container only) — generate tests/golden/*.npz from the UNMODIFIED reference.

    python oracle/gen_golden.py            # regenerate every case
    python oracle/gen_golden.py g20_control

For each case: build a reference-schema input graph (ccbs_b200.scenario.synthetic_input_graph), run
the reference's own scenario generator on it (``Model(network=G)``, simulation/model.py:353), construct
the reference ``CyberBattleCompressedEnv`` behind ``RandomSwitchEnv`` with a seeded default-architecture
GAE encoder, and step it with seeded actions / pre-drawn uniforms / pre-drawn starters
(oracle/trace.py).  The fixture stores the *structure* the reference generator produced (JSON inside
the npz) plus the seeds that regenerate every 768-d vector, and the per-step trace.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import ccbs_b200 as cb                                                  # noqa: E402
from ccbs_b200.gae import GaeWeights                                    # noqa: E402
from ccbs_b200.scenario import spec_to_dict, spec_from_dict, embeddings_of_input_graph  # noqa: E402
from oracle import trace as tr                                          # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> parameters.  `cfg` entries override EnvConfig defaults (= agents/config/train_config.yaml).
CASES = {
    # ~20-node generator-default scenarios (configs[1])
    "g20_control": dict(graph_seed=11, nodes=20, steps=1500, cfg=dict(goal="control")),
    "g12_discovery": dict(graph_seed=12, nodes=12, steps=800, cfg=dict(goal="discovery")),
    "g16_disruption": dict(graph_seed=13, nodes=16, steps=800, cfg=dict(goal="disruption", remove_main_obstacles=False)),
    # 32-node (configs[2])
    "g32_control": dict(graph_seed=14, nodes=32, steps=1500, cfg=dict(goal="control")),
    # long episodes: no proportional cut-off, goal does not stop the episode, clamped reward
    "g10_long": dict(graph_seed=15, nodes=10, steps=900,
                     cfg=dict(goal="control", proportional_cutoff_coefficient=0, episode_iterations=120,
                              stop_at_goal_reached=False, absolute_reward=True, remove_main_obstacles=False)),
    # 100-node "default-like" scenario (configs[0]); >64 nodes exercises 4-word masks
    "g100_control": dict(graph_seed=16, nodes=100, steps=1200, pool_size=560, services_range=(1, 4),
                         vulns_per_service_range=(8, 18), cfg=dict(goal="control")),
    # BASELINE configs[0] at its stated length: 10 000 steps of one default-like 100-node env with random actions.  Oracle-only
    # fixture (the CUDA replay of 10 k single steps through the split calls would take minutes; tests/test_gpu_golden.py skips it,
    # the same scenario class is replayed on the GPU by g100_control)
    "g100_10k": dict(graph_seed=17, nodes=100, steps=10000, pool_size=560, services_range=(1, 4),
                     vulns_per_service_range=(8, 18), cfg=dict(goal="control")),
    # every DoS row filtered out of the action table (compressed:536-538), non-default reward scale, stricter isolation filter
    "g14_removeall": dict(graph_seed=19, nodes=14, steps=600,
                          cfg=dict(goal="discovery", remove_all_obstacles=True, remove_main_obstacles=False, winning_reward=300,
                                   losing_reward=-700, isolation_filter_threshold=0.3, proportional_cutoff_coefficient=2)),
    # node-specific goals (one interest node per env object; observation carries its embedding: 256 + 2 floats)
    "n12_control_node": dict(graph_seed=21, nodes=12, steps=700, interest=5,
                             cfg=dict(goal="control_node", proportional_cutoff_coefficient=2, interest_node_value=200)),
    "n12_discovery_node": dict(graph_seed=22, nodes=12, steps=700, interest=3,
                               cfg=dict(goal="discovery_node", proportional_cutoff_coefficient=3, interest_node_value=150)),
    "n10_disruption_node": dict(graph_seed=23, nodes=10, steps=700, interest=7,
                                cfg=dict(goal="disruption_node", proportional_cutoff_coefficient=3)),
    # scripted attacker (oracle.trace.policy_pick): owns many nodes, escalates, reaches the goal -> winning reward,
    # large action tables, many snapshot slots.  `policy` = noise sigma added to the chosen table row.
    "p6_control_win": dict(graph_seed=30, nodes=6, steps=500, policy=0.02,
                           cfg=dict(goal="control", proportional_cutoff_coefficient=25, episode_iterations=400)),
    "p8_control_win": dict(graph_seed=53, nodes=8, steps=500, policy=0.02,
                           cfg=dict(goal="control", proportional_cutoff_coefficient=25, episode_iterations=400)),
    "p6_control_nostop": dict(graph_seed=30, nodes=6, steps=500, policy=0.02,
                              cfg=dict(goal="control", proportional_cutoff_coefficient=25, episode_iterations=400,
                                       stop_at_goal_reached=False)),
    "p24_control_big": dict(graph_seed=18, nodes=24, steps=700, policy=0.02,
                            cfg=dict(goal="control", proportional_cutoff_coefficient=5, episode_iterations=200)),
    # re-imaging static defender (_env/static_defender.py:27-60): scan draws and detection uniforms are pre-drawn inputs.
    # Random attacker + aggressive defender: the starter is evicted -> owned list empties -> lost.
    "d12_reimage_random": dict(graph_seed=41, nodes=12, steps=600,
                               cfg=dict(goal="control", static_defender_agent="reimage", detect_probability=0.3, scan_capacity=3,
                                        scan_frequency=2, proportional_cutoff_coefficient=4)),
    # scripted attacker that also makes nodes persistent: re-imaged nodes come back owned (env:427-430), get re-owned by
    # lateral moves (duplicates in owned_nodes), stale feature vectors in the visible graph, long episodes
    "d8_reimage_policy": dict(graph_seed=53, nodes=8, steps=900, policy=0.02, p_persist=0.25,
                              cfg=dict(goal="control", static_defender_agent="reimage", detect_probability=0.25, scan_capacity=3,
                                       scan_frequency=3, proportional_cutoff_coefficient=25, episode_iterations=400)),
    "d6_reimage_discovery": dict(graph_seed=30, nodes=6, steps=700, policy=0.02, p_persist=0.3,
                                 cfg=dict(goal="discovery", static_defender_agent="reimage", detect_probability=0.6, scan_capacity=4,
                                          scan_frequency=5, proportional_cutoff_coefficient=30, episode_iterations=300,
                                          stop_at_goal_reached=False)),
    # precise_graph_encoding: every step re-encodes (compressed:455-462)
    "g12_precise": dict(graph_seed=12, nodes=12, steps=500, cfg=dict(goal="control", precise_graph_encoding=True,
                                                                      proportional_cutoff_coefficient=3)),
    # precise_action_space_positions (compressed:419-427,498-506): rows of pairs connected to the action's nodes are
    # refreshed with the current embeddings at every table-maintaining encode
    "g12_positions": dict(graph_seed=12, nodes=12, steps=500, cfg=dict(goal="control", precise_action_space_positions=True,
                                                                        proportional_cutoff_coefficient=3)),
    "p8_positions": dict(graph_seed=53, nodes=8, steps=500, policy=0.02,
                         cfg=dict(goal="control", precise_action_space_positions=True, proportional_cutoff_coefficient=25,
                                  episode_iterations=400)),
    # decode metrics other than cosine (compressed:571-576: np.linalg.norm of action - rows, ord 1 / 2 / inf)
    "g12_l1": dict(graph_seed=12, nodes=12, steps=500, cfg=dict(goal="control", distance_metric="l1", proportional_cutoff_coefficient=3)),
    "g16_l2": dict(graph_seed=13, nodes=16, steps=500, cfg=dict(goal="control", distance_metric="l2", proportional_cutoff_coefficient=3)),
    "g12_inf": dict(graph_seed=12, nodes=12, steps=500, cfg=dict(goal="discovery", distance_metric="inf", proportional_cutoff_coefficient=3)),
    "p8_l2": dict(graph_seed=53, nodes=8, steps=500, policy=0.02,
                  cfg=dict(goal="control", distance_metric="l2", proportional_cutoff_coefficient=25, episode_iterations=400)),
    "p8_inf": dict(graph_seed=53, nodes=8, steps=400, policy=0.02,
                   cfg=dict(goal="control", distance_metric="inf", proportional_cutoff_coefficient=25, episode_iterations=400)),
    # sample_subset_samples (compressed:553-567): at most k rows per outcome class stay in the table after every table-
    # maintaining encode; np.random.choice replaced on both sides by ccbs_b200.philox.subset_keep.  ORACLE-ONLY fixtures (s*):
    # the CUDA path does not implement the sub-sampling yet
    "s16_subset_random": dict(graph_seed=13, nodes=16, steps=600, philox_seed=77,
                              cfg=dict(goal="control", sample_subset_samples=12, proportional_cutoff_coefficient=3)),
    "s8_subset_policy": dict(graph_seed=53, nodes=8, steps=500, policy=0.02, philox_seed=78,
                             cfg=dict(goal="control", sample_subset_samples=25, proportional_cutoff_coefficient=25,
                                      episode_iterations=400)),
    "s12_subset_positions": dict(graph_seed=12, nodes=12, steps=400, philox_seed=79,
                                 cfg=dict(goal="control", sample_subset_samples=10, precise_action_space_positions=True,
                                          proportional_cutoff_coefficient=3)),
    # ExternalRandomEvents static defender (_env/static_defender.py:63-161): services stopped / started, firewall rules added /
    # removed at random; every step re-encodes (compressed:401) over stale node features.  ORACLE-ONLY fixtures (e*)
    "e12_events_random": dict(graph_seed=41, nodes=12, steps=600,
                              cfg=dict(goal="control", static_defender_agent="events", random_event_probability=0.02,
                                       proportional_cutoff_coefficient=4)),
    "e8_events_policy": dict(graph_seed=53, nodes=8, steps=600, policy=0.02,
                             cfg=dict(goal="control", static_defender_agent="events", random_event_probability=0.03,
                                      proportional_cutoff_coefficient=25, episode_iterations=300)),
    # precise_action_space_positions under the re-imaging defender: rows refreshed around `changed_nodes` = the nodes re-imaged in
    # this step (compressed:423-427).  ORACLE-ONLY.  (With the events defender the reference raises NodeNotFound, see EnvConfig.)
    "x8_reimage_positions": dict(graph_seed=53, nodes=8, steps=500, policy=0.02, p_persist=0.25,
                                 cfg=dict(goal="control", static_defender_agent="reimage", detect_probability=0.25, scan_capacity=3,
                                          scan_frequency=3, precise_action_space_positions=True, proportional_cutoff_coefficient=25,
                                          episode_iterations=300)),
    "p6_l1": dict(graph_seed=30, nodes=6, steps=400, policy=0.02,
                  cfg=dict(goal="control", distance_metric="l1", proportional_cutoff_coefficient=25, episode_iterations=400)),
}
POOL_SEED = 1234
GAE_SEED = 0


def build_case(name):
    p = CASES[name]
    from oracle import ref_bridge as rb
    pool = cb.synthetic_vuln_pool(POOL_SEED, p.get("pool_size", 200))
    gkw = {k: p[k] for k in ("services_range", "vulns_per_service_range") if k in p}
    graph = cb.synthetic_input_graph(p["graph_seed"], p["nodes"], pool=pool, **gkw)
    model = rb.reference_model_from_input_graph(graph, seed=p["graph_seed"])
    spec = cb.spec_from_model(model, name=name)
    cfg = cb.EnvConfig(**p["cfg"])
    weights = GaeWeights.random(GAE_SEED)
    return p, graph, model, spec, cfg, weights


def make_case_inputs(p):
    """Random-action cases: actions U(-4,4)^905.  Policy cases: `actions` is N(0, sigma) noise added to a table row."""
    actions, uniforms = tr.make_inputs(p["graph_seed"] * 1000 + 1, p["steps"])
    if "policy" in p:
        rng = np.random.default_rng(p["graph_seed"] * 1000 + 4)
        actions = (rng.standard_normal(actions.shape) * p["policy"]).astype(np.float32)
    return actions, uniforms


def make_case_defender_draws(p, cfg):
    if cfg.static_defender_agent is None:
        return None
    if cfg.static_defender_agent == "events":
        return tr.make_events_draws(p["graph_seed"] * 1000 + 5, p["steps"], p["nodes"], float(cfg.random_event_probability))
    return tr.make_defender_draws(p["graph_seed"] * 1000 + 5, p["steps"], p["nodes"], int(cfg.scan_capacity))


def generate(name):
    from oracle import ref_bridge as rb
    t0 = time.time()
    p, graph, model, spec, cfg, weights = build_case(name)
    interest = p.get("interest")
    from oracle.cbs_oracle import OracleEnv
    cb.compile_scenarios([spec], cfg.isolation_filter_threshold)               # cross-checks reach counts against the Model
    feasible = OracleEnv(spec, weights, cfg, interest_node=interest).feasible_starters()
    assert feasible, "no feasible starter for this case"
    actions, uniforms = make_case_inputs(p)
    starters = tr.make_starters(p["graph_seed"] * 1000 + 2, feasible, p["steps"] + 2)
    runner = rb.ReferenceRunner(model, weights, cfg, interest_node=interest, subset_vuln_index=tr.vuln_index(spec),
                                philox_seed=p.get("philox_seed", 0))
    rec = tr.record(tr.ReferenceAdapter(runner, spec), actions, uniforms, starters,
                    policy_seed=(p["graph_seed"] * 1000 + 3) if "policy" in p else None,
                    defender_draws=make_case_defender_draws(p, cfg), p_persist=p.get("p_persist", 0.0))
    meta = dict(name=name, params={k: v for k, v in p.items() if k != "cfg"}, cfg=p["cfg"], pool_seed=POOL_SEED,
                gae_seed=GAE_SEED, input_seed=p["graph_seed"] * 1000 + 1, starter_seed=p["graph_seed"] * 1000 + 2,
                spec=spec_to_dict(spec),
                emb_checksum=float(sum(float(np.sum(v)) for v in spec.vuln_emb.values())))
    out = {k: v for k, v in rec.items()}
    # store obs compactly: unique consecutive rows + index
    obs = out.pop("obs")
    change = np.ones(len(obs), bool)
    change[1:] = np.any(obs[1:] != obs[:-1], axis=1)
    out["obs_rows"] = obs[change]
    out["obs_idx"] = (np.cumsum(change) - 1).astype(np.int32)
    out["starters"] = starters[: int(rec["num_episodes"])]
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    eps = int(rec["num_episodes"])
    print(f"{name}: {p['steps']} steps, {eps} episodes, {os.path.getsize(path) / 1e6:.2f} MB, {time.time() - t0:.1f}s; "
          f"success codes seen: {sorted(set(rec['code'].tolist()))}; end reasons: {np.bincount(rec['reason'], minlength=4).tolist()}; "
          f"max owned {int((rec['owned_order'] >= 0).sum(1).max())}, max discovered {int((rec['disc_order'] >= 0).sum(1).max())}; "
          f"re-imaged per episode {rec['stats'][:, 9].tolist() if len(rec['stats']) else []}")


def load_case(path):
    """Rebuild (spec, cfg, weights, actions, uniforms, starters, trace) from a fixture — no reference needed."""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    p = meta["params"]
    pool = cb.synthetic_vuln_pool(meta["pool_seed"], p.get("pool_size", 200))
    gkw = {k: tuple(p[k]) for k in ("services_range", "vulns_per_service_range") if k in p}
    graph = cb.synthetic_input_graph(p["graph_seed"], p["nodes"], pool=pool, **gkw)
    vuln_emb, service_fv = embeddings_of_input_graph(graph)
    spec = spec_from_dict(meta["spec"], vuln_emb, service_fv)
    chk = float(sum(float(np.sum(v)) for v in spec.vuln_emb.values()))
    if abs(chk - meta["emb_checksum"]) > 1e-6 * max(1.0, abs(chk)):
        raise RuntimeError("golden fixture: regenerated embeddings do not match the recorded checksum")
    cfg = cb.EnvConfig(**meta["cfg"])
    weights = GaeWeights.random(meta["gae_seed"])
    actions, uniforms = make_case_inputs(p)
    rec = {k: z[k] for k in z.files if k not in ("meta", "obs_rows", "obs_idx")}
    rec["obs"] = z["obs_rows"][z["obs_idx"]]
    return dict(meta=meta, spec=spec, cfg=cfg, weights=weights, actions=actions, uniforms=uniforms,
                starters=z["starters"], trace=rec, policy_seed=0 if "policy" in p else None,
                policy_rows=rec.get("policy_rows"), interest=p.get("interest"), philox_seed=p.get("philox_seed", 0),
                defender_draws=make_case_defender_draws(p, cfg))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        generate(n)
