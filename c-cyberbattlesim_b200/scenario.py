"""Scenario model and scenario compiler.

A *scenario* is the immutable part of one attack graph: nodes, services, firewall rules,
vulnerabilities with their predicted outcomes, and the reachability counts the reset needs.
Three ways to obtain one:

* :func:`spec_from_model` — duck-typed extraction from a reference ``Model`` (or a pickled one):
  reads exactly the fields the hot path reads (simulation/model.py:212-338).  No import of the
  reference is needed; any object with those attributes works.
* :func:`synthetic_input_graph` + :func:`spec_from_input_graph` — a seeded generator of input
  graphs in the reference's *input* schema (simulation/generate_network.py:144-220) and a
  restatement of how ``cyberbattle_model_from_nodes_graph`` (generate_network.py:97-312) turns it
  into nodes (numpy RNG instead of the global ``random`` — scenario randomisation is not on the
  hot path, so the draws need not match, the *rules* do).
* :func:`synthetic_spec` — both of the above in one call.

:func:`compile_scenarios` flattens a list of specs into the structure-of-arrays tables the CUDA
kernels consume (one global, de-duplicated vulnerability-embedding table shared by all scenarios).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import constants as C


# --------------------------------------------------------------------------------------------
# Spec (plain Python, order-preserving — order is semantically relevant)
# --------------------------------------------------------------------------------------------
@dataclass
class ResultSpec:
    """One PredictedResult (model.py:204-210)."""
    kind: int                     # constants.K_*
    vtype: int                    # 0 local / 1 remote (VulnerabilityType)
    nodes: List[int] = field(default_factory=list)   # Reconnaissance.nodes as node indices (ordered)
    level: int = C.PRIV_ROOT      # PrivilegeEscalation.level


@dataclass
class VulnSpec:
    """One VulnerabilityInfo on one node (model.py:212-248)."""
    vid: str
    port: int
    priv_required: int            # 0 / 1 / 3
    success_rate: float
    cost: float
    results: List[ResultSpec]


@dataclass
class ServiceSpec:
    port: int
    running: bool
    fv: np.ndarray                # float64[768] service feature vector


@dataclass
class NodeSpec:
    """One NodeInfo (model.py:294-338) — initial values."""
    node_id: str
    tag: str = ""
    value: int = 0
    has_data: bool = False
    visible: bool = True
    level_at_access: int = C.PRIV_USER
    reimageable: bool = True
    sla_weight: float = 1.0
    services: List[ServiceSpec] = field(default_factory=list)
    fw_in: List[Tuple[int, int]] = field(default_factory=list)    # (port, 0 ALLOW / 1 BLOCK), rule order
    fw_out: List[Tuple[int, int]] = field(default_factory=list)
    vulns: List[VulnSpec] = field(default_factory=list)           # dict order of NodeInfo.vulnerabilities


@dataclass
class ScenarioSpec:
    nodes: List[NodeSpec]
    vuln_emb: Dict[str, np.ndarray]     # vulnerability_ID -> float64[768] (env.vulnerabilities_embeddings)
    name: str = ""
    # optional: reach counts taken from a reference Model (used only to cross-check the compiler)
    ref_counts: Optional[Dict[str, np.ndarray]] = None

    @property
    def num_nodes(self) -> int:
        return len(self.nodes)


# --------------------------------------------------------------------------------------------
# Extraction from a reference Model (duck-typed)
# --------------------------------------------------------------------------------------------
_KIND_BY_CLASSNAME = {
    "DenialOfService": C.K_DOS, "Discovery": C.K_DISCOVERY, "Collection": C.K_COLLECTION,
    "Exfiltration": C.K_EXFILTRATION, "Reconnaissance": C.K_RECON, "DefenseEvasion": C.K_EVASION,
    "Persistence": C.K_PERSISTENCE, "PrivilegeEscalation": C.K_PRIVESC,
    "CredentialAccess": C.K_CREDACCESS, "LateralMove": C.K_LATERAL, "Execution": C.K_EXECUTION,
}


def _vector(v, feature_extractor):
    """A feature vector is a list of floats, or a dict {LM name: list} before Model.update_feature_vectors
    (model.py:423-430) resolved it."""
    if isinstance(v, dict):
        if feature_extractor is None or feature_extractor not in v:
            raise ValueError(f"feature vectors are per-LM dicts {sorted(v)}; pass feature_extractor=<one of them>")
        v = v[feature_extractor]
    return np.asarray(v, dtype=np.float64)


def spec_from_model(model, name: str = "", feature_extractor: Optional[str] = None) -> ScenarioSpec:
    """Flatten a reference ``Model`` (model.py:349-437).  Feature vectors may already be resolved to one LM
    (``update_feature_vectors``, model.py:423-430) or still be per-LM dicts (then ``feature_extractor`` picks one).
    Node order = order of ``model.network.nodes`` (that order defines starter indices and graph insertion order)."""
    net = model.network
    ids = list(net.nodes)
    index = {nid: i for i, nid in enumerate(ids)}
    nodes: List[NodeSpec] = []
    vuln_emb: Dict[str, np.ndarray] = {}
    for nid in ids:
        info = net.nodes[nid]["data"]
        services = [ServiceSpec(port=int(s.name), running=bool(s.running),
                                fv=_vector(s.feature_vector, feature_extractor)) for s in info.services]
        vulns = []
        for vid, v in info.vulnerabilities.items():
            results = []
            for r in v.results:
                kind = _KIND_BY_CLASSNAME.get(type(r.outcome).__name__)
                if kind is None:
                    continue
                rs = ResultSpec(kind=kind, vtype=int(r.type.value))
                if kind == C.K_RECON:
                    rs.nodes = [index[n] for n in r.outcome.nodes]
                if kind == C.K_PRIVESC:
                    rs.level = int(r.outcome.level)
                results.append(rs)
            vulns.append(VulnSpec(vid=str(vid), port=int(v.port), priv_required=int(v.privileges_required),
                                  success_rate=float(v.rates.successRate), cost=float(v.cost), results=results))
            # create_vulnerabilities_embeddings (compressed:614-618): later nodes overwrite earlier ones
            vuln_emb[str(vid)] = _vector(v.embedding, feature_extractor)
        nodes.append(NodeSpec(
            node_id=str(nid), tag=str(info.tag), value=int(info.value), has_data=bool(info.has_data),
            visible=bool(info.visible), level_at_access=int(info.level_at_access),
            reimageable=bool(info.reimageable), sla_weight=float(info.sla_weight), services=services,
            fw_in=[(int(r.port), int(r.permission.value)) for r in info.firewall.incoming],
            fw_out=[(int(r.port), int(r.permission.value)) for r in info.firewall.outgoing],
            vulns=vulns))
    spec = ScenarioSpec(nodes=nodes, vuln_emb=vuln_emb, name=name)
    # reference reach counts for cross-checking (cyberbattle_env.py:205-217)
    try:
        counts = {}
        for key, paths in (("ownable", model.access_shortest_paths), ("discoverable", model.knows_shortest_paths),
                           ("disruptable", model.dos_shortest_paths)):
            counts[key] = np.array([sum(1 for k, v in paths[nid].items() if k != nid and v is not None)
                                    for nid in ids], dtype=np.int32)
        spec.ref_counts = counts
    except Exception:
        spec.ref_counts = None
    return spec


def load_scenario_folder(folder: str, nlp_extractor: str, pca_components: int = 768, subset: Optional[str] = None):
    """Load the reference's scenario layout ``<folder>/<id>/network_<nlp>.pkl`` (+ ``pca/num_components=<k>/`` when PCA was
    applied) exactly as agents/train_agent.py:380-390 does, optionally restricted to the ids of ``split.yaml``'s
    ``training_set`` / ``validation_set`` (agents/train_agent.py:419-428).  The pickles hold reference ``Model`` objects, so
    the reference package must be importable for unpickling; nothing else of it is used.  Returns (ids, specs)."""
    import os
    import pickle
    ids = sorted(int(e) for e in os.listdir(folder) if e.isdigit() and os.path.isdir(os.path.join(folder, e)))
    if subset is not None:
        import yaml
        with open(os.path.join(folder, "split.yaml")) as f:
            wanted = {int(e["id"]) for e in yaml.safe_load(f)[subset]}
        ids = [i for i in ids if i in wanted]
    specs = []
    for i in ids:
        sub = os.path.join(folder, str(i))
        if pca_components != 768:
            sub = os.path.join(sub, "pca", f"num_components={pca_components}")
        with open(os.path.join(sub, f"network_{nlp_extractor}.pkl"), "rb") as f:
            model = pickle.load(f)
        specs.append(spec_from_model(model, name=str(i), feature_extractor=nlp_extractor))
    return ids, specs


# --------------------------------------------------------------------------------------------
# Synthetic input graphs (reference input schema) and their conversion
# --------------------------------------------------------------------------------------------
_CLASS_LABELS = ["reconnaissance", "discovery", "persistence", "credential access", "collection",
                 "privilege escalation", "DOS", "lateral move", "execution", "defense evasion",
                 "exfiltration"]
# outcome mix loosely following docs/ch2_env_stats.md:38-49 (lateral move / recon frequent)
_CLASS_WEIGHTS = np.array([0.16, 0.08, 0.07, 0.06, 0.09, 0.10, 0.10, 0.16, 0.04, 0.07, 0.07])


def synthetic_vuln_pool(seed: int, pool_size: int = 200, dim: int = C.VULN_EMB_DIM):
    """A pool of vulnerability records in the reference's scraped-DB schema
    (generate_network.py:154-220 reads exactly these keys)."""
    rng = np.random.default_rng(seed)
    pool = []
    for k in range(pool_size):
        n_cls = int(rng.integers(1, 5))
        labels = list(rng.choice(len(_CLASS_LABELS), size=n_cls, replace=False, p=_CLASS_WEIGHTS / _CLASS_WEIGHTS.sum()))
        pool.append({
            "ID": f"CVE-SYN-{seed}-{k:04d}",
            "description": "",
            "attack_vector": str(rng.choice(["NETWORK", "ADJACENT_NETWORK", "LOCAL", "PHYSICAL"], p=[0.6, 0.1, 0.25, 0.05])),
            "attack_complexity": str(rng.choice(["LOW", "MEDIUM", "HIGH"], p=[0.6, 0.25, 0.15])),
            "privileges_required": str(rng.choice(["NONE", "LOW", "HIGH"], p=[0.6, 0.3, 0.1])),
            "user_interaction": "NONE",
            "confidentiality_impact": str(rng.choice(["NONE", "PARTIAL", "COMPLETE"], p=[0.2, 0.4, 0.4])),
            "integrity_impact": "PARTIAL",
            "availability_impact": "PARTIAL",
            "base_score": float(np.round(rng.uniform(2, 10), 1)),
            "exploitability_score": float(np.round(rng.uniform(1, 10), 1)),
            "impact_score": float(np.round(rng.uniform(1, 10), 1)),
            "base_severity": "MEDIUM",
            "classes": [{"class": _CLASS_LABELS[i], "probability": float(np.round(rng.uniform(0.2, 1.0), 3))}
                        for i in labels],
            "feature_vector": rng.standard_normal(dim).astype(np.float32).astype(np.float64),
        })
    return pool


def synthetic_input_graph(seed: int, num_nodes: int, pool=None, pool_seed: int = 1234, pool_size: int = 200,
                          services_range=(1, 3), vulns_per_service_range=(3, 12), dim: int = C.VULN_EMB_DIM):
    """Input graph as ``{node_id: {"category":..., "services":[...]}}`` in the schema
    ``cyberbattle_model_from_nodes_graph`` consumes (generate_network.py:144-220)."""
    rng = np.random.default_rng(seed)
    if pool is None:
        pool = synthetic_vuln_pool(pool_seed, pool_size, dim)
    graph = {}
    for n in range(num_nodes):
        n_services = int(rng.integers(services_range[0], services_range[1] + 1))
        ports = rng.choice(np.array([21, 22, 23, 25, 53, 80, 110, 139, 443, 445, 3306, 3389, 5432, 8080]),
                           size=n_services, replace=False)
        services = []
        for p in ports:
            n_v = int(rng.integers(vulns_per_service_range[0], vulns_per_service_range[1] + 1))
            picks = rng.choice(len(pool), size=min(n_v, len(pool)), replace=False)
            services.append({
                "port": int(p), "product": f"svc{int(p)}", "version": "1.0", "description": "",
                "feature_vector": rng.standard_normal(dim).astype(np.float32).astype(np.float64),
                "vulnerabilities": [pool[int(i)] for i in picks],
            })
        graph[f"n{n}"] = {"category": str(rng.choice(["windows", "unix", "iot"])), "services": services}
    return graph


def _scale_prob(rng, prob_range, score, score_range=(0, 10)):
    """generate_network.py:19-30 scale_probability_range_with_score."""
    if score is None:
        return float(rng.uniform(prob_range[0], prob_range[1]))
    if score < 0:
        return 0.0
    return prob_range[0] + ((score - score_range[0]) / (score_range[1] - score_range[0])) * (prob_range[1] - prob_range[0])


_CONF = {None: 0, "NONE": -1, "PARTIAL": 0.5, "LOW": 0.5, "COMPLETE": 1, "HIGH": 1}     # generate_network.py:46-57
_COMPLEXITY = {None: 0, "LOW": 1, "MEDIUM": 0.75, "HIGH": 0.5}                            # :60-68
_VTYPE = {"LOCAL": 0, "PHYSICAL": 0, "NETWORK": 1, "ADJACENT_NETWORK": 1}                 # :71-81
_PRIV = {None: 1, "NONE": 0, "LOW": 1, "SINGLE": 1, "HIGH": 3, "MULTIPLE": 3}             # :84-94


def spec_from_input_graph(graph: dict, seed: int, *, firewall_rule_incoming_probability=0.2,
                          firewall_rule_outgoing_probability=0.2, knows_neighbor_probability_range=(0.2, 0.3),
                          data_presence_probability=0.5, partial_visibility_probability=0.5,
                          need_to_escalate_probability=0.5, service_shutdown_probability=0.1,
                          success_rate_probability_range=(0.9, 1.0), value_range=(0, 100), name="") -> ScenarioSpec:
    """Restates generate_network.py:97-256 (node construction + a-posteriori firewall rules)."""
    rng = np.random.default_rng(seed)
    ids = list(graph.keys())
    index = {nid: i for i, nid in enumerate(ids)}
    nodes: List[NodeSpec] = []
    vuln_emb: Dict[str, np.ndarray] = {}
    for nid in ids:
        services, vulns_by_id, order = [], {}, []
        has_data, partially_visible, level_at_access = False, False, C.PRIV_ROOT
        node_ports = []
        for svc in graph[nid]["services"]:
            node_ports.append(int(svc["port"]))
            shutdown = rng.random() < service_shutdown_probability
            fv = svc.get("feature_vector")
            fv = np.zeros(C.VULN_EMB_DIM) if fv is None else np.asarray(fv, dtype=np.float64)
            services.append(ServiceSpec(port=int(svc["port"]), running=not shutdown, fv=fv))
            for v in svc.get("vulnerabilities", []):
                vtype = _VTYPE[v["attack_vector"]]
                results, discovery_list = [], []
                for cls in v["classes"]:
                    kind = C.LABEL_TO_KIND.get(cls["class"])
                    if kind is None:
                        continue
                    rs = ResultSpec(kind=kind, vtype=vtype)
                    if kind == C.K_COLLECTION:
                        if rng.random() < data_presence_probability:
                            has_data = True
                    elif kind == C.K_RECON:
                        p = _scale_prob(rng, knows_neighbor_probability_range, _CONF[v["confidentiality_impact"]])
                        for other in ids:
                            if other != nid and rng.random() < p:
                                discovery_list.append(index[other])
                        # Reconnaissance(nodes) keeps a reference to discovery_list when it is non-empty
                        rs.nodes = discovery_list if discovery_list else []
                    elif kind == C.K_DISCOVERY:
                        if rng.random() < partial_visibility_probability:
                            partially_visible = True
                    elif kind == C.K_PRIVESC:
                        if rng.random() < need_to_escalate_probability:
                            level_at_access = C.PRIV_USER
                    results.append(rs)
                spec_v = VulnSpec(vid=str(v["ID"]), port=int(svc["port"]), priv_required=_PRIV[v["privileges_required"]],
                                  success_rate=float(_scale_prob(rng, success_rate_probability_range,
                                                                 _COMPLEXITY[v["attack_complexity"]])),
                                  cost=float(10 - v["exploitability_score"]), results=results)
                if spec_v.vid not in vulns_by_id:      # dict semantics: overwrite keeps first position (:214)
                    order.append(spec_v.vid)
                vulns_by_id[spec_v.vid] = spec_v
                vuln_emb[spec_v.vid] = np.asarray(v["feature_vector"], dtype=np.float64)
        for vs in vulns_by_id.values():               # freeze shared recon lists
            for r in vs.results:
                r.nodes = list(r.nodes)
        nodes.append(NodeSpec(
            node_id=str(nid), tag=str(graph[nid].get("category", "")),
            value=int(rng.integers(value_range[0], value_range[1] + 1)), has_data=has_data,
            visible=not partially_visible, level_at_access=level_at_access, services=services,
            fw_in=[(p, 0) for p in node_ports], fw_out=[(p, 0) for p in node_ports],
            vulns=[vulns_by_id[k] for k in order]))
    for nd in nodes:                                   # generate_network.py:242-256
        nd.fw_in = [(p, 1 if rng.random() < firewall_rule_incoming_probability else 0) for p, _ in nd.fw_in]
        nd.fw_out = [(p, 1 if rng.random() < firewall_rule_outgoing_probability else 0) for p, _ in nd.fw_out]
    return ScenarioSpec(nodes=nodes, vuln_emb=vuln_emb, name=name)


def synthetic_spec(seed: int, num_nodes: int, pool=None, pool_seed: int = 1234, pool_size: int = 200, **kw) -> ScenarioSpec:
    gkw = {k: kw.pop(k) for k in ("services_range", "vulns_per_service_range") if k in kw}
    g = synthetic_input_graph(seed, num_nodes, pool=pool, pool_seed=pool_seed, pool_size=pool_size, **gkw)
    return spec_from_input_graph(g, seed + 7919, name=f"syn{seed}_n{num_nodes}", **kw)


# --------------------------------------------------------------------------------------------
# Compiler: specs -> SoA tables
# --------------------------------------------------------------------------------------------
def check_events_compatible(spec: "ScenarioSpec") -> None:
    """The device keeps the ExternalRandomEvents defender's mutable state (static_defender_actions.py:96-168) as one bit per
    (node, service slot): that needs distinct service ports per node, at most 16 services, and firewall rules on the node's own
    service ports only (what the reference's generator produces, generate_network.py:222-256)."""
    for nd in spec.nodes:
        ports = [s.port for s in nd.services]
        if len(ports) > 16 or len(set(ports)) != len(ports):
            raise ValueError(f"scenario '{spec.name}', node {nd.node_id}: the events defender needs <= 16 services with distinct ports")
        if len(ports) == 0:
            raise ValueError(f"scenario '{spec.name}', node {nd.node_id}: a node without services makes the reference's "
                             "ExternalRandomEvents raise (random.choice of an empty list, static_defender.py:123,150)")
        for rules in (nd.fw_in, nd.fw_out):
            rp = [p for p, _ in rules]
            if any(p not in ports for p in rp) or len(set(rp)) != len(rp):
                raise ValueError(f"scenario '{spec.name}', node {nd.node_id}: the events defender needs one firewall rule per own service port")


def _is_passing(rules: Sequence[Tuple[int, int]], port: int) -> bool:
    """attacker_actions.py:550-559 — first rule on the port decides, default allow."""
    for p, perm in rules:
        if p == port:
            return perm == 0
    return True


def _reach_counts(adj: np.ndarray) -> np.ndarray:
    """#nodes reachable from each node (excluding itself) — equals the count of non-None entries of
    the all-pairs shortest-path row with the source popped (cyberbattle_env.py:205-217)."""
    n = adj.shape[0]
    reach = adj.copy() | np.eye(n, dtype=bool)
    for k in range(n):                                # Warshall closure on bool rows
        reach |= np.outer(reach[:, k], reach[k, :])
    np.fill_diagonal(reach, False)
    return reach.sum(axis=1).astype(np.int32), reach


def scenario_graphs(spec: ScenarioSpec):
    """knows / access / dos adjacency (generate_network.py:258-306)."""
    n = spec.num_nodes
    knows = np.zeros((n, n), dtype=bool)
    for i, nd in enumerate(spec.nodes):
        for v in nd.vulns:
            for r in v.results:
                if r.kind == C.K_RECON:
                    for j in r.nodes:
                        if j != i:
                            knows[i, j] = True
    _, knows_reach = _reach_counts(knows)
    access = np.zeros((n, n), dtype=bool)
    dos = np.zeros((n, n), dtype=bool)
    for i, nd in enumerate(spec.nodes):
        for v in nd.vulns:
            for r in v.results:
                if r.kind not in (C.K_LATERAL, C.K_CREDACCESS, C.K_DOS):
                    continue
                if any(p == v.port and perm == 1 for p, perm in nd.fw_in):
                    continue
                for s, snd in enumerate(spec.nodes):
                    if s == i or not knows_reach[s, i]:
                        continue
                    if any(p == v.port and perm == 1 for p, perm in snd.fw_out):
                        continue
                    if r.kind == C.K_DOS:
                        dos[s, i] = True
                    else:
                        access[s, i] = True
    return knows, access, dos


@dataclass
class ScenarioTables:
    """Flattened immutable tables for S scenarios (numpy, host side).  Index spaces:
    node rows are global (``node_off[s] + j``), vulnerability instances are global
    (``inst``), candidate rows are global (``row``), recon list entries are global."""
    num_scenarios: int
    max_nodes: int                 # max nodes over scenarios
    words: int                     # mask words = ceil(max_nodes / 32)
    # per scenario
    sc_num_nodes: np.ndarray       # i32[S]
    sc_node_off: np.ndarray        # i32[S+1]
    sc_num_ports: np.ndarray       # i32[S]
    sc_port_off: np.ndarray        # i32[S+1]    into outblock
    sc_num_uvuln: np.ndarray       # i32[S]      unique vulnerability ids in the scenario
    sc_uvuln_off: np.ndarray       # i32[S+1]    into uvuln_global / inst_of
    sc_instof_off: np.ndarray      # i64[S+1]    into inst_of (N_s * U_s entries per scenario)
    sc_discoverable_amount: np.ndarray  # i32[S]
    sc_init_has_data: np.ndarray   # u32[S, words]
    sc_init_visible: np.ndarray    # u32[S, words]
    sc_feasible_off: np.ndarray    # i32[6, S+1]  per goal (constants.GOAL_*): offsets into feasible_starters[goal]
    feasible_starters: List[np.ndarray]   # 6 x i32[...]  starters passing the isolation filter / able to reach the interest node
    sc_interest: np.ndarray        # i32[S]  interest node per scenario (-1: none; *_node goals need one)
    # per node (global node index)
    nd_value: np.ndarray           # i32
    nd_level_at_access: np.ndarray  # u8
    nd_reimageable: np.ndarray     # u8    NodeInfo.reimageable (model.py:312), read by the static defender
    nd_ownable: np.ndarray         # i32   reach counts with this node as starter
    nd_discoverable: np.ndarray    # i32
    nd_disruptable: np.ndarray     # i32
    nd_row_off: np.ndarray         # i32[Nn+1, 2]-> flattened [2*Nn+1]: local list then remote list per node
    # per port of a scenario: nodes whose outgoing firewall blocks it
    outblock: np.ndarray           # u32[sum ports, words]
    # ExternalRandomEvents defender: per scenario node the initial { running services, incoming BLOCK, outgoing BLOCK } bit sets
    # over the node's service slots and its service count; per instance the target's service slot of the vulnerability's port;
    # per (scenario port, node) the node's service slot of that port (0xFF = the node has no service on it)
    nd_ev_init: np.ndarray         # u16[Nn, 4]
    vi_svc_slot: np.ndarray        # u8[I]
    out_slot: np.ndarray           # u8[sum ports, max_nodes]
    # per unique vulnerability of a scenario
    uvuln_global: np.ndarray       # i32   row in the global embedding table
    inst_of: np.ndarray            # i32   [node j][u] -> instance index or -1, per scenario block
    # per vulnerability instance
    vi_port: np.ndarray            # i32   scenario-local port index
    vi_flags: np.ndarray           # u32   see VI_* below
    vi_kinds_any: np.ndarray       # u16   outcome kinds present (any type)      -> local exploit  (:397-400)
    vi_kinds_remote: np.ndarray    # u16   outcome kinds with a REMOTE result    -> remote exploit (:149-152)
    vi_success: np.ndarray         # f64
    vi_cost: np.ndarray            # f64
    vi_recon_any: np.ndarray       # i32[.,2] (off, len) of first Recon result, any type
    vi_recon_remote: np.ndarray    # i32[.,2] (off, len) of first REMOTE Recon result
    vi_ulocal: np.ndarray          # i32   scenario-local unique vuln index
    recon_nodes: np.ndarray        # u8
    # candidate rows (action-table templates, compressed:621-637)
    row_packed: np.ndarray         # u32   global_vuln | kind<<20 | onehot<<24
    row_inst: np.ndarray           # i32   instance index
    # global vulnerability embedding table
    vemb64: np.ndarray             # f64[Ug, 768]
    vemb32: np.ndarray             # f32[Ug, 768]
    vnorm2: np.ndarray             # f64[Ug]  ||v||^2 (without the one-hot 1)
    # bookkeeping for host-side naming
    node_ids: List[List[str]]
    vuln_ids: List[List[str]]      # per scenario, scenario-local unique index -> id
    global_vuln_ids: List[str]
    specs: List[ScenarioSpec] = field(default_factory=list, repr=False)


# vi_flags bits
VI_LISTENING = 1 << 0      # vuln.port among the node's running services (attacker_actions.py:161)
VI_IN_ALLOWED = 1 << 1     # target incoming firewall passes the port (:180)
VI_PRIVREQ_SHIFT = 2       # 2 bits: privileges_required 0/1/3
VI_LEVEL_ANY_SHIFT = 4     # 2 bits: PrivilegeEscalation.level of first privesc result (any type)
VI_LEVEL_REMOTE_SHIFT = 6  # 2 bits: same, first REMOTE privesc result


def compile_scenarios(specs: Sequence[ScenarioSpec], isolation_filter_threshold: float = 0.1,
                      check_ref_counts: bool = True, interest_nodes: Optional[Sequence[int]] = None,
                      interest_node_value: Optional[int] = None) -> ScenarioTables:
    """``interest_nodes[s]`` (one node index per scenario) enables the *_node goals: it defines their feasible starters
    (cyberbattle_env.py:249-275) and, with ``interest_node_value``, overrides that node's value (:277) in the compiled
    tables and in the specs kept for the GAE folding."""
    S = len(specs)
    if interest_nodes is not None:
        if len(interest_nodes) != S:
            raise ValueError("interest_nodes needs one entry per scenario")
        patched = []
        for spec, it in zip(specs, interest_nodes):
            if not (0 <= int(it) < spec.num_nodes):
                raise ValueError(f"interest node {it} out of range for scenario '{spec.name}'")
            if interest_node_value is not None:
                nodes = list(spec.nodes)
                nodes[int(it)] = dataclasses.replace(nodes[int(it)], value=int(interest_node_value))
                spec = dataclasses.replace(spec, nodes=nodes)
            patched.append(spec)
        specs = patched
    max_nodes = max(s.num_nodes for s in specs)
    if max_nodes > C.MAX_NODES:
        raise ValueError(f"scenario with {max_nodes} nodes exceeds MAX_NODES={C.MAX_NODES}")
    words = (max_nodes + 31) // 32

    # global embedding table, de-duplicated by (id, bytes)
    gkey: Dict[Tuple[str, bytes], int] = {}
    gemb: List[np.ndarray] = []
    gids: List[str] = []

    sc_num_nodes, sc_node_off = [], [0]
    sc_num_ports, sc_port_off = [], [0]
    sc_num_uvuln, sc_uvuln_off, sc_instof_off = [], [0], [0]
    sc_da, sc_hd, sc_vis = [], [], []
    feas = [[] for _ in range(6)]
    feas_off = [[0] for _ in range(6)]
    nd_value, nd_laa, nd_own, nd_disc, nd_disr, nd_reim = [], [], [], [], [], []
    nd_ev_init, vi_svc_slot, out_slot = [], [], []       # ExternalRandomEvents defender (mutable services / firewall rules)
    nd_row_off = [0]
    outblock = []
    uvuln_global, inst_of = [], []
    vi_port, vi_flags, vi_ka, vi_kr, vi_succ, vi_cost, vi_ra, vi_rr, vi_ul = [], [], [], [], [], [], [], [], []
    recon_nodes: List[int] = []
    row_packed, row_inst = [], []
    node_ids, vuln_ids = [], []

    for spec in specs:
        n = spec.num_nodes
        sc_num_nodes.append(n)
        node_ids.append([nd.node_id for nd in spec.nodes])
        # ports
        ports: Dict[int, int] = {}
        for nd in spec.nodes:
            for v in nd.vulns:
                ports.setdefault(v.port, len(ports))
        ob = np.zeros((max(len(ports), 1), words), dtype=np.uint32)
        for p, pi in ports.items():
            for j, nd in enumerate(spec.nodes):
                if not _is_passing(nd.fw_out, p):
                    ob[pi, j // 32] |= np.uint32(1 << (j % 32))
        outblock.append(ob)
        # slot of port p among node j's services (first match, like get_service_index cyberbattle_env.py:547-551), 0xFF = none
        osl = np.full((max(len(ports), 1), max_nodes), 0xFF, dtype=np.uint8)
        for p, pi in ports.items():
            for j, nd in enumerate(spec.nodes):
                sp = [s.port for s in nd.services]
                if p in sp and sp.index(p) < 16:
                    osl[pi, j] = sp.index(p)
        out_slot.append(osl)
        sc_num_ports.append(ob.shape[0])
        sc_port_off.append(sc_port_off[-1] + ob.shape[0])
        # unique vulnerability ids of the scenario
        uloc: Dict[str, int] = {}
        for nd in spec.nodes:
            for v in nd.vulns:
                if v.vid not in uloc:
                    uloc[v.vid] = len(uloc)
                    emb = np.ascontiguousarray(spec.vuln_emb[v.vid], dtype=np.float64)
                    if emb.shape != (C.VULN_EMB_DIM,):
                        raise ValueError(f"vulnerability {v.vid}: embedding shape {emb.shape}, expected ({C.VULN_EMB_DIM},)")
                    k = (v.vid, emb.tobytes())
                    if k not in gkey:
                        gkey[k] = len(gemb)
                        gemb.append(emb)
                        gids.append(v.vid)
                    uvuln_global.append(gkey[k])
        U = len(uloc)
        if len(gemb) >= (1 << 20):
            raise ValueError("global vulnerability table exceeds 2^20 rows")
        vuln_ids.append(list(uloc.keys()))
        sc_num_uvuln.append(U)
        sc_uvuln_off.append(sc_uvuln_off[-1] + U)
        instof = np.full((n, max(U, 1)), -1, dtype=np.int32)
        ubase = sc_uvuln_off[-2]
        # nodes, instances, candidate rows
        hd = np.zeros(words, dtype=np.uint32)
        vis = np.zeros(words, dtype=np.uint32)
        da = n
        for j, nd in enumerate(spec.nodes):
            if nd.has_data:
                hd[j // 32] |= np.uint32(1 << (j % 32))
                da += 2
            if nd.visible:
                vis[j // 32] |= np.uint32(1 << (j % 32))
            else:
                da += 1
            nd_value.append(nd.value)
            nd_laa.append(nd.level_at_access)
            nd_reim.append(1 if nd.reimageable else 0)
            running_ports = [s.port for s in nd.services if s.running]
            svc_ports = [s.port for s in nd.services]
            run_bits = sum(1 << i for i, s in enumerate(nd.services[:16]) if s.running)
            in_bits = sum(1 << i for i, p in enumerate(svc_ports[:16]) if not _is_passing(nd.fw_in, p))
            out_bits = sum(1 << i for i, p in enumerate(svc_ports[:16]) if not _is_passing(nd.fw_out, p))
            nd_ev_init.append((run_bits, in_bits, out_bits, len(nd.services)))
            local_rows, remote_rows = [], []
            for v in nd.vulns:
                inst = len(vi_port)
                ul = uloc[v.vid]
                instof[j, ul] = inst
                flags = 0
                if v.port in running_ports:
                    flags |= VI_LISTENING
                if _is_passing(nd.fw_in, v.port):
                    flags |= VI_IN_ALLOWED
                flags |= (v.priv_required & 3) << VI_PRIVREQ_SHIFT
                ka = kr = 0
                ra = rr = None
                la = lr = None
                for r in v.results:
                    if not (ka >> r.kind) & 1:
                        ka |= 1 << r.kind
                        if r.kind == C.K_RECON:
                            ra = r.nodes
                        if r.kind == C.K_PRIVESC:
                            la = r.level
                    if r.vtype == 1 and not (kr >> r.kind) & 1:
                        kr |= 1 << r.kind
                        if r.kind == C.K_RECON:
                            rr = r.nodes
                        if r.kind == C.K_PRIVESC:
                            lr = r.level
                    oh = C.onehot_index(r.vtype, r.kind)
                    if oh is not None:
                        packed = uvuln_global[ubase + ul] | (r.kind << 20) | (oh << 24)
                        (remote_rows if r.vtype == 1 else local_rows).append((packed, inst))
                flags |= ((la if la is not None else C.PRIV_ROOT) & 3) << VI_LEVEL_ANY_SHIFT
                flags |= ((lr if lr is not None else C.PRIV_ROOT) & 3) << VI_LEVEL_REMOTE_SHIFT
                for lst, dst in ((ra, vi_ra), (rr, vi_rr)):
                    if lst is None:
                        dst.append((0, 0))
                    else:
                        dst.append((len(recon_nodes), len(lst)))
                        recon_nodes.extend(int(x) for x in lst)
                vi_port.append(ports[v.port])
                vi_svc_slot.append(svc_ports.index(v.port) if v.port in svc_ports and svc_ports.index(v.port) < 16 else 0xFF)
                vi_flags.append(flags)
                vi_ka.append(ka)
                vi_kr.append(kr)
                vi_succ.append(v.success_rate)
                vi_cost.append(v.cost)
                vi_ul.append(ul)
            for lst in (local_rows, remote_rows):
                for packed, inst in lst:
                    row_packed.append(packed)
                    row_inst.append(inst)
                nd_row_off.append(len(row_packed))
        inst_of.append(instof.reshape(-1))
        sc_instof_off.append(sc_instof_off[-1] + instof.size)
        sc_da.append(da)
        sc_hd.append(hd)
        sc_vis.append(vis)
        # reachability (cyberbattle_env.py:205-217) and feasible starters (:219-248)
        knows, access, dos = scenario_graphs(spec)
        own, own_reach = _reach_counts(access)
        disc, disc_reach = _reach_counts(knows)
        disr, disr_reach = _reach_counts(dos)
        if check_ref_counts and spec.ref_counts is not None:
            for key, mine in (("ownable", own), ("discoverable", disc), ("disruptable", disr)):
                if not np.array_equal(spec.ref_counts[key], mine):
                    raise AssertionError(f"{spec.name}: compiled {key} counts differ from the reference Model's")
        nd_own.extend(own.tolist())
        nd_disc.extend(disc.tolist())
        nd_disr.extend(disr.tolist())
        thr = isolation_filter_threshold * n
        for g, cnt in ((C.GOAL_CONTROL, own), (C.GOAL_DISCOVERY, disc), (C.GOAL_DISRUPTION, disr)):
            ok = [j for j in range(n) if not (cnt[j] < thr)]
            feas[g].extend(ok)
            feas_off[g].append(len(feas[g]))
        it = -1 if interest_nodes is None else int(interest_nodes[len(sc_node_off) - 1])
        for g, reach in ((C.GOAL_CONTROL_NODE, own_reach), (C.GOAL_DISCOVERY_NODE, disc_reach), (C.GOAL_DISRUPTION_NODE, disr_reach)):
            ok = [] if it < 0 else [j for j in range(n) if j != it and reach[j, it]]     # cyberbattle_env.py:249-275
            feas[g].extend(ok)
            feas_off[g].append(len(feas[g]))
        sc_node_off.append(sc_node_off[-1] + n)

    vemb64 = np.stack(gemb).astype(np.float64) if gemb else np.zeros((1, C.VULN_EMB_DIM))
    return ScenarioTables(
        num_scenarios=S, max_nodes=max_nodes, words=words,
        sc_num_nodes=np.array(sc_num_nodes, np.int32), sc_node_off=np.array(sc_node_off, np.int32),
        sc_num_ports=np.array(sc_num_ports, np.int32), sc_port_off=np.array(sc_port_off, np.int32),
        sc_num_uvuln=np.array(sc_num_uvuln, np.int32), sc_uvuln_off=np.array(sc_uvuln_off, np.int32),
        sc_instof_off=np.array(sc_instof_off, np.int64),
        sc_discoverable_amount=np.array(sc_da, np.int32),
        sc_init_has_data=np.stack(sc_hd), sc_init_visible=np.stack(sc_vis),
        sc_feasible_off=np.array(feas_off, np.int32),
        feasible_starters=[np.array(f, np.int32) for f in feas],
        sc_interest=np.array([-1] * S if interest_nodes is None else [int(x) for x in interest_nodes], np.int32),
        nd_value=np.array(nd_value, np.int32), nd_level_at_access=np.array(nd_laa, np.uint8),
        nd_reimageable=np.array(nd_reim, np.uint8),
        nd_ownable=np.array(nd_own, np.int32), nd_discoverable=np.array(nd_disc, np.int32),
        nd_disruptable=np.array(nd_disr, np.int32), nd_row_off=np.array(nd_row_off, np.int32),
        outblock=np.concatenate(outblock, axis=0),
        nd_ev_init=np.array(nd_ev_init, np.uint16).reshape(-1, 4), vi_svc_slot=np.array(vi_svc_slot, np.uint8),
        out_slot=np.concatenate(out_slot, axis=0),
        uvuln_global=np.array(uvuln_global, np.int32), inst_of=np.concatenate(inst_of),
        vi_port=np.array(vi_port, np.int32), vi_flags=np.array(vi_flags, np.uint32),
        vi_kinds_any=np.array(vi_ka, np.uint16), vi_kinds_remote=np.array(vi_kr, np.uint16),
        vi_success=np.array(vi_succ, np.float64), vi_cost=np.array(vi_cost, np.float64),
        vi_recon_any=np.array(vi_ra, np.int32).reshape(-1, 2), vi_recon_remote=np.array(vi_rr, np.int32).reshape(-1, 2),
        vi_ulocal=np.array(vi_ul, np.int32),
        recon_nodes=np.array(recon_nodes, np.uint8),
        row_packed=np.array(row_packed, np.uint32), row_inst=np.array(row_inst, np.int32),
        vemb64=vemb64, vemb32=vemb64.astype(np.float32), vnorm2=(vemb64 * vemb64).sum(axis=1),
        node_ids=node_ids, vuln_ids=vuln_ids, global_vuln_ids=gids, specs=list(specs))


# --------------------------------------------------------------------------------------------
# (De)serialisation of the structural part of a spec (embeddings travel separately or by seed)
# --------------------------------------------------------------------------------------------
def spec_to_dict(spec: ScenarioSpec) -> dict:
    """Structure only (no 768-d vectors): JSON-serialisable."""
    return {
        "name": spec.name,
        "nodes": [{
            "id": nd.node_id, "tag": nd.tag, "value": nd.value, "has_data": nd.has_data, "visible": nd.visible,
            "level_at_access": nd.level_at_access, "reimageable": nd.reimageable, "sla_weight": nd.sla_weight,
            "services": [[s.port, bool(s.running)] for s in nd.services],
            "fw_in": [list(r) for r in nd.fw_in], "fw_out": [list(r) for r in nd.fw_out],
            "vulns": [{"id": v.vid, "port": v.port, "priv": v.priv_required, "succ": v.success_rate, "cost": v.cost,
                       "res": [[r.kind, r.vtype, list(r.nodes), r.level] for r in v.results]} for v in nd.vulns],
        } for nd in spec.nodes],
        "ref_counts": None if spec.ref_counts is None else {k: [int(x) for x in v] for k, v in spec.ref_counts.items()},
    }


def spec_from_dict(d: dict, vuln_emb: Dict[str, np.ndarray], service_fv: Dict[Tuple[str, int], np.ndarray]) -> ScenarioSpec:
    """Inverse of :func:`spec_to_dict`; ``service_fv[(node_id, port)]`` supplies service vectors."""
    nodes = []
    for nd in d["nodes"]:
        nodes.append(NodeSpec(
            node_id=nd["id"], tag=nd["tag"], value=int(nd["value"]), has_data=bool(nd["has_data"]),
            visible=bool(nd["visible"]), level_at_access=int(nd["level_at_access"]), reimageable=bool(nd["reimageable"]),
            sla_weight=float(nd["sla_weight"]),
            services=[ServiceSpec(port=int(p), running=bool(r), fv=np.asarray(service_fv[(nd["id"], int(p))], dtype=np.float64))
                      for p, r in nd["services"]],
            fw_in=[(int(p), int(q)) for p, q in nd["fw_in"]], fw_out=[(int(p), int(q)) for p, q in nd["fw_out"]],
            vulns=[VulnSpec(vid=v["id"], port=int(v["port"]), priv_required=int(v["priv"]), success_rate=float(v["succ"]),
                            cost=float(v["cost"]),
                            results=[ResultSpec(kind=int(k), vtype=int(ty), nodes=[int(x) for x in ns], level=int(lv))
                                     for k, ty, ns, lv in v["res"]]) for v in nd["vulns"]]))
    used = {v.vid for nd in nodes for v in nd.vulns}
    spec = ScenarioSpec(nodes=nodes, vuln_emb={k: np.asarray(vuln_emb[k], dtype=np.float64) for k in used}, name=d.get("name", ""))
    if d.get("ref_counts"):
        spec.ref_counts = {k: np.array(v, dtype=np.int32) for k, v in d["ref_counts"].items()}
    return spec


def embeddings_of_input_graph(graph: dict):
    """(vuln_emb, service_fv) dictionaries of a reference-schema input graph."""
    vuln_emb, service_fv = {}, {}
    for nid, nd in graph.items():
        for svc in nd["services"]:
            service_fv[(str(nid), int(svc["port"]))] = np.asarray(svc["feature_vector"], dtype=np.float64)
            for v in svc.get("vulnerabilities", []):
                vuln_emb[str(v["ID"])] = np.asarray(v["feature_vector"], dtype=np.float64)
    return vuln_emb, service_fv
