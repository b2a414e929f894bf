"""TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's continuous-env step path.

This file restates, in plain Python/numpy (+ torch fp32 for the graph encoder), what
C-CyberBattleSim computes on ``RandomSwitchEnv.step/reset -> CyberBattleCompressedEnv.step/reset``
for one environment.  Every method cites the reference file:line it follows.  It is **not** part of
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` leg may import it.  The product path (``c-cyberbattlesim_b200/``) never does.

Parity status: PINNED against the reference's own Python code, run unmodified in the build
container under the stub modules in ``oracle/shims`` by ``oracle/gen_golden.py``; the recorded traces
live in ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them through this file.  The
graph-encoder layers (torch_geometric 2.5.3 NNConv / GCNConv / BatchNorm) are third-party code that is
absent from /root/reference and from this image; their published semantics are restated both here
and in the shim, so the GAE part is pinned to that restatement, not to a PyG binary.

Inputs are the product's own neutral containers (``ScenarioSpec``, ``GaeWeights``, ``EnvConfig``) —
*not* the compiled SoA tables — so a GPU-vs-oracle test also covers the scenario compiler.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial import distance as _sp_distance

import ccbs_b200.constants as C

_LOCAL_LABELS = [C.K_DOS, C.K_DISCOVERY, C.K_COLLECTION, C.K_EXFILTRATION, C.K_RECON, C.K_EVASION,
                 C.K_PERSISTENCE, C.K_PRIVESC]                                  # compressed:595-597
_REMOTE_LABELS = [C.K_DOS, C.K_DISCOVERY, C.K_COLLECTION, C.K_EXFILTRATION, C.K_RECON, C.K_EVASION,
                  C.K_PERSISTENCE, C.K_CREDACCESS, C.K_LATERAL]                 # compressed:598-600
_REENCODE_KINDS = (C.K_LATERAL, C.K_DOS, C.K_RECON)                             # compressed:462
_REFRESH_KINDS = (C.K_DISCOVERY, C.K_COLLECTION, C.K_PERSISTENCE, C.K_PRIVESC, C.K_EXFILTRATION,
                  C.K_EVASION, C.K_DOS, C.K_LATERAL, C.K_CREDACCESS)            # compressed:472-479


class _Node:
    """Mutable NodeInfo fields (model.py:294-338) over an immutable NodeSpec."""
    __slots__ = ("spec", "agent_installed", "privilege_level", "status", "has_data", "data_collected",
                 "data_exfiltrated", "visible", "persistence", "defense_evasion", "vulns", "fw_in", "fw_out", "running")

    def __init__(self, spec):
        self.spec = spec
        self.agent_installed = False
        self.privilege_level = C.PRIV_NONE
        self.status = C.ST_RUNNING
        self.has_data = spec.has_data
        self.data_collected = False
        self.data_exfiltrated = False
        self.visible = spec.visible
        self.persistence = False
        self.defense_evasion = False
        self.vulns = {v.vid: v for v in spec.vulns}
        # mutable only under the ExternalRandomEvents defender (static_defender_actions.py:96-168)
        self.fw_in = [(int(p), int(q)) for p, q in spec.fw_in]
        self.fw_out = [(int(p), int(q)) for p, q in spec.fw_out]
        self.running = [bool(sv.running) for sv in spec.services]


class GaeOracle:
    """gae/model.py:70-82 GAEEncoder.forward for the default layer config, eval mode, torch fp32 on CPU.
    NNConv / GCNConv / BatchNorm follow torch_geometric 2.5.3 (see module docstring)."""

    def __init__(self, w):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
        self.nn0_w, self.nn0_b = t(w.nn0_w), t(w.nn0_b)
        self.nn2_w, self.nn2_b = t(w.nn2_w), t(w.nn2_b)
        self.root_w, self.conv1_b = t(w.root_w), t(w.conv1_b)
        self.gcn_w, self.gcn_b = t(w.gcn_w), t(w.gcn_b)
        self.bn1 = {k: t(v) for k, v in w.bn1.items()}
        self.bn2 = {k: t(v) for k, v in w.bn2.items()}
        self.eps = w.bn_eps

    def _bn(self, x, bn):
        return torch.nn.functional.batch_norm(x, bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"],
                                              False, 0.0, self.eps)

    @torch.no_grad()
    def forward(self, x, edge_index, edge_attr):
        n, d_in, d_out = x.shape[0], x.shape[1], self.conv1_b.shape[0]
        src, dst = edge_index[0], edge_index[1]
        # NNConv (aggr=add, root weight, bias)
        h = torch.relu(torch.nn.functional.linear(edge_attr, self.nn0_w, self.nn0_b))
        theta = torch.nn.functional.linear(h, self.nn2_w, self.nn2_b).view(-1, d_in, d_out)
        msg = torch.matmul(x[src].unsqueeze(1), theta).squeeze(1)
        out = torch.zeros(n, d_out, dtype=x.dtype).index_add_(0, dst, msg)
        out = out + torch.nn.functional.linear(x, self.root_w) + self.conv1_b
        out = torch.relu(self._bn(out, self.bn1))
        # GCNConv: add_remaining_self_loops, symmetric normalisation by in-degree
        keep = src != dst
        loop = torch.arange(n, dtype=src.dtype)
        s2, d2 = torch.cat([src[keep], loop]), torch.cat([dst[keep], loop])
        deg = torch.zeros(n, dtype=x.dtype).index_add_(0, d2, torch.ones(s2.shape[0], dtype=x.dtype))
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        hw = torch.nn.functional.linear(out, self.gcn_w)
        agg = torch.zeros_like(hw).index_add_(0, d2, (dinv[s2] * dinv[d2]).unsqueeze(1) * hw[s2])
        return torch.relu(self._bn(agg + self.gcn_b, self.bn2))


class OracleEnv:
    """One continuous env.  ``step(action, uniform)`` = RandomSwitchEnv.step -> CyberBattleCompressedEnv.step
    with the success-rate draw replaced by the supplied uniform (consumed only where the reference
    calls ``random.random()``, attacker_actions.py:190,409)."""

    def __init__(self, spec, gae_weights, cfg, interest_node=None, philox_seed=0, env_index=0):
        self.spec, self.cfg = spec, cfg
        # sample_subset_samples (compressed:83,105,521-522,553-567): rows kept per outcome class, 0 = off.  The reference draws
        # the subset with np.random.choice; here (and in the golden generator, which hands the same function to the reference)
        # it is ccbs_b200.philox.subset_keep keyed by (philox_seed, env_index, lifetime balance counter, row identity).
        self.sample_subset_samples = int(getattr(cfg, "sample_subset_samples", 0) or 0)
        self.philox_seed, self.env_index, self.balance_calls, self.rows_dropped = int(philox_seed), int(env_index), 0, 0
        self._vloc = {}
        for nd in spec.nodes:
            for v in nd.vulns:
                self._vloc.setdefault(v.vid, len(self._vloc))
        self.node_goal = cfg.goal.endswith("node")
        self.interest = None if interest_node is None else int(interest_node)      # cyberbattle_env.py:127-131 (fixed per env object)
        if self.node_goal and self.interest is None:
            raise ValueError("*_node goals need an interest node")
        self.gae = GaeOracle(gae_weights)
        self.N = spec.num_nodes
        self.rew, self.pen = dict(cfg.rewards_dict), dict(cfg.penalties_dict)
        self.pen.setdefault("node_already_stopped", self.pen.get("machine_already_stopped", -10))
        self.goal = cfg.goal
        self.episode_iterations = cfg.episode_iterations
        # ScanAndReimageCompromisedMachines (_env/static_defender.py:27-60); None -> no defender
        self.defender = getattr(cfg, "static_defender_agent", None) == "reimage"
        # ExternalRandomEvents (_env/static_defender.py:63-161)
        self.events_defender = getattr(cfg, "static_defender_agent", None) == "events"
        self.always_encode = self.defender or self.events_defender or \
            bool(getattr(cfg, "precise_graph_encoding", False))                                     # compressed:401,455-462
        self.distance_metric = getattr(cfg, "distance_metric", "cosine")                            # compressed:82,100
        self.precise_positions = bool(getattr(cfg, "precise_action_space_positions", False))        # compressed:86,419-427
        self.proportional_cutoff_coefficient = cfg.proportional_cutoff_coefficient
        # create_vulnerabilities_embeddings (compressed:614-618)
        self.vuln_emb = {k: np.asarray(v, dtype=np.float64) for k, v in spec.vuln_emb.items()}
        # create_vulnerabilities_embeddings_per_node_type (compressed:621-637)
        self.per_node_type = []
        for nd in spec.nodes:
            lists = {0: [], 1: []}
            for v in nd.vulns:
                for ri, r in enumerate(v.results):
                    oh = self._onehot(r.vtype, r.kind)
                    if oh is None:
                        continue
                    lists[r.vtype].append((v.vid, ri, r.kind, np.concatenate((self.vuln_emb[v.vid], oh))))
            self.per_node_type.append(lists)
        # reach counts: all-pairs shortest paths of the access / knows / dos graphs
        # (model.py:418-420 + networkx_utils.py:19-58), generate_network.py:258-306 for the graphs
        self._reach = self._reach_counts()
        self.episode_id = 0
        self.done = False
        self.nodes = None

    # ---- static helpers -------------------------------------------------------------------
    @staticmethod
    def _onehot(vtype, kind):
        """compressed:593-611 map_outcome_to_onehot"""
        if kind == C.K_EXECUTION:
            return None
        labels = _LOCAL_LABELS if vtype == 0 else _REMOTE_LABELS
        if kind not in labels:
            return None
        oh = np.zeros(C.OUTCOME_DIM, dtype=np.float64)
        oh[labels.index(kind)] = 1
        return oh

    def _reach_counts(self):
        N, nodes = self.N, self.spec.nodes
        knows = [set() for _ in range(N)]
        for i, nd in enumerate(nodes):                               # generate_network.py:259-267
            for v in nd.vulns:
                for r in v.results:
                    if r.kind == C.K_RECON:
                        knows[i].update(j for j in r.nodes if j != i)

        def reach_from(adj, s):
            seen, stack = {s}, [s]
            while stack:
                u = stack.pop()
                for w in adj[u]:
                    if w not in seen:
                        seen.add(w)
                        stack.append(w)
            seen.discard(s)
            return seen
        knows_reach = [reach_from(knows, s) for s in range(N)]
        access = [set() for _ in range(N)]
        dos = [set() for _ in range(N)]
        for i, nd in enumerate(nodes):                               # generate_network.py:271-306
            for v in nd.vulns:
                for r in v.results:
                    if r.kind not in (C.K_LATERAL, C.K_CREDACCESS, C.K_DOS):
                        continue
                    if any(p == v.port and perm == 1 for p, perm in nd.fw_in):
                        continue
                    for s in range(N):
                        if s == i or i not in knows_reach[s]:        # nx.has_path(knows_graph, s, i)
                            continue
                        if any(p == v.port and perm == 1 for p, perm in nodes[s].fw_out):
                            continue
                        (dos if r.kind == C.K_DOS else access)[s].add(i)
        self._reach_sets = {"control": [reach_from(access, s) for s in range(N)], "discovery": knows_reach,
                            "disruption": [reach_from(dos, s) for s in range(N)]}
        return {k: [len(x) for x in v] for k, v in self._reach_sets.items()}

    def feasible_starters(self):
        if self.node_goal:                                           # cyberbattle_env.py:249-275
            base = self.goal[:-5]
            return [s for s in range(self.N) if s != self.interest and self.interest in self._reach_sets[base][s]]
        thr = self.cfg.isolation_filter_threshold * self.N           # cyberbattle_env.py:206
        return [s for s in range(self.N) if not (self._reach[self.goal][s] < thr)]

    # ---- reset ----------------------------------------------------------------------------
    def reset(self, starter=None, rng=None):
        """compressed:158-189 reset -> cyberbattle_env.py:134-186 reset_env (+ pick_starter_node :189-296)."""
        self.nodes = [_Node(nd) for nd in self.spec.nodes]           # deepcopy of the pristine network (:145)
        if starter is None:
            if not self.cfg.random_starter_node:
                starter = 0                                          # :191
            else:
                feas = self.feasible_starters()
                if not feas:
                    raise RuntimeError("NoSuitableStarterNode")      # :198-199
                # rejection sampling with randrange(N) (:201) == uniform over feasible starters
                starter = feas[int(rng.integers(len(feas)))] if rng is not None else feas[0]
        self.starter = int(starter)
        self.ownable_count = self._reach["control"][self.starter]            # :209
        self.discoverable_count = self._reach["discovery"][self.starter]     # :213
        self.disruptable_count = self._reach["disruption"][self.starter]     # :217
        self.proportional_nodes = {"control": self.ownable_count, "discovery": self.discoverable_count,
                                   "disruption": self.disruptable_count}[self.goal[:-5] if self.node_goal else self.goal]   # :226-275
        if self.node_goal:
            self.node_value = {self.interest: self.cfg.interest_node_value}      # :277 set_node_property(interest, "value", ...)
        else:
            self.node_value = {}
        # :279-288 — both dict-key tests are true for every node, so every node counts
        self.discoverable_amount = 0
        for nd in self.nodes:
            self.discoverable_amount += 1
            if nd.has_data:
                self.discoverable_amount += 2
            if not nd.visible:
                self.discoverable_amount += 1
        self.nodes[self.starter].agent_installed = True              # :296
        self.network_availability = 1.0
        self.episode_id += 1
        self.done = False
        self.truncated = False
        self.end_episode_reason = 0
        self.num_iterations = 0
        self.discovered_amount = 0
        self.stepcount = 0
        self.discovered_nodes = [self.starter]                       # :178
        self.owned_nodes = [self.starter]                            # :179
        # AttackerAgentActions.__init__ (attacker_actions.py:50-53): starter owned at level_at_access
        self._discovered = {}                                        # node -> last_owned (bool) tracking
        self._stale = set()                                          # nodes re-imaged after they were last owned
        self.reimaging = {}                                          # StaticDefenderAgentActions.node_reimaging_progress (:23)
        self.changed_nodes = []
        self.overall_reimaged, self.num_events = [], 0               # cyberbattle_env.py:160-161
        self._mark_owned(self.starter, self.nodes[self.starter].spec.level_at_access)
        # compressed:167-188
        self.processed_pairs = set()
        self.graph_nodes = []            # insertion order of the evolving visible graph
        self.node_x = {}
        self.graph_edges = {}            # src -> {tgt: float64[768]}
        self._add_graph_node(self.starter)
        self.action_keys, self.action_rows = [], []
        self._row_of_key = {}            # the reference's table is a dict: re-adding a key overwrites its row in place
        self._rows_cache = None
        self.exploited = {}
        self.node_embeddings, emb = self.encode()
        self.create_continuous_action_space()
        self.observation = {"graph_embeddings": emb, "discrete_features": self._discrete_features()}
        self.n_encodes = 1
        return self.observation

    # ---- attacker actions (simulation/attacker_actions.py) -----------------------------------
    def _mark_owned(self, n, level):
        """attacker_actions.py:70-89 __mark_node_as_owned; returns (was ever owned, is currently owned)."""
        ever = self._discovered.get(n)
        was = bool(ever)
        self._discovered.setdefault(n, False)
        node = self.nodes[n]
        node.agent_installed = True
        node.privilege_level = max(int(node.privilege_level), int(level))    # model.py:340
        self._discovered[n] = True
        cur = was and n not in self._stale                                   # attacker_actions.py:561-573
        self._stale.discard(n)                                               # last_owned_at = now (:87)
        return was, cur

    def _value(self, n):
        return self.node_value.get(n, self.nodes[n].spec.value)

    @staticmethod
    def _passing(rules, port):
        """attacker_actions.py:550-559"""
        for p, perm in rules:
            if p == port:
                return perm == 0
        return True

    def _exploit_remote(self, s, t, vid, kind, u):
        """attacker_actions.py:92-351.  Returns (reward, obtained code, recon list or None)."""
        P, R = self.pen, self.rew
        src, tgt = self.nodes[s], self.nodes[t]
        if not src.agent_installed:
            return P["invalid_action"], C.OC_INVALID_SRC_NOT_OWNED, None                  # :109
        if t not in self._discovered:
            return P["invalid_action"], C.OC_INVALID_TGT_NOT_DISCOVERED, None             # :115
        if src.status != C.ST_RUNNING:
            return P["invalid_action"], C.OC_SRC_NOT_RUNNING, None                        # :121
        if tgt.status != C.ST_RUNNING:
            return P["invalid_action"], C.OC_TGT_NOT_RUNNING, None                        # :127
        v = tgt.vulns.get(vid)
        if v is None:
            return P["no_vulnerability_in_node"], C.OC_NO_VULNERABILITY, None             # :133
        if v.priv_required and not tgt.privilege_level >= v.priv_required:
            return P["no_enough_privileges"], C.OC_NO_PRIVILEGE, None                     # :141
        res = next((r for r in v.results if r.vtype == 1 and r.kind == kind), None)       # :149-152
        if res is None:
            return P["invalid_action"], C.OC_OUTCOME_NOT_PRESENT, None                    # :153
        if v.port not in [sv.port for sv, up in zip(tgt.spec.services, tgt.running) if up]:
            return P["scanning_unopen_port"], C.OC_PORT_NOT_LISTENING, None               # :161
        if not src.defense_evasion and not self._passing(src.fw_out, v.port):
            return P["blocked_by_local_firewall"], C.OC_FW_OUTGOING, None                 # :170
        if not tgt.defense_evasion and not self._passing(tgt.fw_in, v.port):
            return P["blocked_by_remote_firewall"], C.OC_FW_INCOMING, None                # :180
        if u >= v.success_rate:
            return P["success_rate_failed"], C.OC_UNSUCCESSFUL, None                      # :190
        return self._apply(t, tgt, v, res, remote=True)

    def _exploit_local(self, n, vid, kind, u):
        """attacker_actions.py:353-547"""
        P = self.pen
        node = self.nodes[n]
        if not node.agent_installed:
            return P["invalid_action"], C.OC_INVALID_SRC_NOT_OWNED, None                  # :363
        if node.status != C.ST_RUNNING:
            return P["invalid_action"], C.OC_SRC_NOT_RUNNING, None                        # :370
        v = node.vulns.get(vid)
        if v is None:
            return P["no_vulnerability_in_node"], C.OC_NO_VULNERABILITY, None             # :377
        if v.priv_required and not node.privilege_level >= v.priv_required:
            return P["no_enough_privileges"], C.OC_NO_PRIVILEGE, None                     # :386
        res = next((r for r in v.results if r.kind == kind), None)                        # :397-400 (no type test)
        if res is None:
            return P["invalid_action"], C.OC_OUTCOME_NOT_PRESENT, None                    # :401
        if u >= v.success_rate:
            return P["success_rate_failed"], C.OC_UNSUCCESSFUL, None                      # :409
        return self._apply(n, node, v, res, remote=False)

    def _apply(self, t, tgt, v, res, remote):
        """Per-outcome mutation and reward: attacker_actions.py:199-351 (remote) / :418-547 (local)."""
        P, R = self.pen, self.rew
        total, recon = 0, None
        k = res.kind
        if k == C.K_COLLECTION:
            if not tgt.has_data:
                return P["no_data_to_collect"], C.OC_NO_NEEDED, None
            tgt.has_data, tgt.data_collected = False, True
            total += R["data_collected_reward"]
        elif k == C.K_PERSISTENCE:
            if tgt.persistence:
                return P["already_persistent"], C.OC_REPEATED, None
            tgt.persistence = True
            total += R["persistence_reward"]
        elif k == C.K_DOS:
            if remote and tgt.status == C.ST_STOPPED:                                     # :232 (unreachable, :127)
                return P["node_already_stopped"], C.OC_REPEATED, None
            tgt.status = C.ST_STOPPED
            total += R["dos_coefficient"] * self._value(t)
        elif k == C.K_DISCOVERY:
            if tgt.visible:
                return P["node_already_visible"], C.OC_REPEATED, None
            tgt.visible = True
            total += R["acquired_visibility_reward"]
        elif k == C.K_EXFILTRATION:
            if not (tgt.data_collected and not tgt.data_exfiltrated):
                return P["no_data_to_exfiltrate"], C.OC_NO_NEEDED, None
            tgt.data_exfiltrated = True
            total += R["data_exfiltrated_reward"]
        elif k == C.K_EVASION:
            if tgt.defense_evasion:
                return P["already_defense_evasion"], C.OC_REPEATED, None
            tgt.defense_evasion = True
            total += R["defense_evaded_reward"]
        elif k == C.K_RECON:
            new = 0
            for n in res.nodes:                                                           # :291-296 / :506-511
                if n not in self._discovered:
                    self._discovered[n] = False
                    new += 1
            total += R["node_discovered_coefficient"] * new
            recon = list(res.nodes)
        elif k == C.K_PRIVESC:
            if remote and tgt.privilege_level == C.PRIV_NONE:                             # :303 (remote only)
                return P["privilege_escalation_in_node_not_owned"], C.OC_NO_PRIVILEGE, None
            if tgt.privilege_level >= res.level:                                          # :309 / :518
                return P["privilege_escalation_to_level_already_had"], C.OC_REPEATED, None
            self._mark_owned(t, res.level)
            total += R["privilege_escalation_reward"]
        elif remote and k in (C.K_LATERAL, C.K_CREDACCESS):
            ever, already = self._mark_owned(t, tgt.spec.level_at_access)                 # :327 (before the test!)
            if already:
                return P["node_already_owned"], C.OC_REPEATED, None                       # :331-335
            if not ever:
                total += R["value_coefficient"] * float(self._value(t))                   # :336-340
        else:
            return P["outcome_not_valid"], C.OC_REMOTE_OUTCOME_LOCAL, None                # :345 / :536-540
        total -= R["cost_coefficient"] * v.cost                                           # :348 / :544
        return total, k, recon

    # ---- episode logic (cyberbattle_env.py) ------------------------------------------------
    def static_defender_step(self, draws):
        """cyberbattle_env.py:416-430 static_defender_step with ScanAndReimageCompromisedMachines.  ``draws`` =
        (node indices for random.choices, uniforms for numpy.random.random() in CALL order)."""
        # StaticDefenderAgentActions.on_attacker_step_taken (static_defender_actions.py:54-68)
        for n in list(self.reimaging):
            if self.reimaging[n] > 0:
                self.reimaging[n] -= 1
            else:
                nd = self.nodes[n]
                nd.status = C.ST_RUNNING
                if nd.persistence:
                    nd.agent_installed = True
                del self.reimaging[n]
        # ScanAndReimageCompromisedMachines.step (static_defender.py:45-60)
        changed = []
        if self.stepcount % int(self.cfg.scan_frequency) == 0:
            scan_nodes, scan_uniforms = draws
            ui = 0
            for n in list(scan_nodes)[: int(self.cfg.scan_capacity)]:
                nd = self.nodes[int(n)]
                if nd.status == C.ST_RUNNING and nd.agent_installed and not nd.defense_evasion:
                    u = float(scan_uniforms[ui])
                    ui += 1
                    if u <= self.cfg.detect_probability and nd.spec.reimageable:
                        self.reimaging[int(n)] = C.REIMAGING_DURATION             # reimage_node (:37-52)
                        nd.agent_installed = False
                        nd.status = C.ST_IMAGING
                        self._stale.add(int(n))                                  # last_reimaging = now
                        changed.append(int(n))
        self.changed_nodes = changed                                              # :418
        self.num_events += len(changed)                                           # :419
        self.overall_reimaged.extend(changed)                                     # :422
        for n in changed:
            self.owned_nodes.remove(n)                                            # :425 (first occurrence)
        for n in range(self.N):                                                   # :427-430 persistence re-owns a re-imaged node
            nd = self.nodes[n]
            if nd.agent_installed and nd.status == C.ST_RUNNING and n not in self.owned_nodes:
                self.owned_nodes.append(n)

    def events_defender_step(self, draws):
        """cyberbattle_env.py:416-419,431 static_defender_step with ExternalRandomEvents (_env/static_defender.py:76-161).
        ``draws[n]`` = (f, u_event, u_pick, u_side) for node n: f indexes random.choice(["start service", "firewall remove",
        "stop service", "firewall add"]) (:80), u_event is the numpy.random.random() compared with the probability, u_pick
        picks the service / port (random.choice -> floor(u_pick * len)), u_side <= 0.5 means an incoming rule.  A node with
        defense evasion consumes only f (:100,:112,:124,:149)."""
        draws = draws[0] if isinstance(draws, tuple) else draws
        events = 0
        self.changed_nodes = []                                                            # nodes_changed (:78,:93-94)
        p = float(self.cfg.random_event_probability)
        for n in range(self.N):                                                            # environment.get_nodes() order
            f, u_event, u_pick, u_side = int(draws[n][0]), float(draws[n][1]), float(draws[n][2]), float(draws[n][3])
            nd = self.nodes[n]
            if nd.defense_evasion or not u_event <= p:
                continue
            services = nd.spec.services
            if f in (0, 2):                                                                # start / stop a service (:98-119)
                if len(services) == 0:
                    continue
                port = services[min(int(u_pick * len(services)), len(services) - 1)].port
                if nd.status == C.ST_RUNNING:                                              # static_defender_actions.py:150,161
                    for i, sv in enumerate(services):
                        if sv.port == port:
                            nd.running[i] = (f == 0)
                events += 1                                                                # counted even when the node is not Running
                self.changed_nodes.append(n)
            else:                                                                          # firewall remove (ALLOW) / add (BLOCK) (:122-161)
                if len(services) == 0:
                    raise IndexError("Cannot choose from an empty sequence")              # random.choice([]) in the reference
                port = services[min(int(u_pick * len(services)), len(services) - 1)].port
                perm = 1 if f == 3 else 0
                if (port, perm) in nd.fw_in:                                               # BOTH sides test firewall.incoming (:135,138 / :158,161)
                    continue
                rules = nd.fw_in if u_side <= 0.5 else nd.fw_out
                patched, seen = [], False                                                  # override_firewall_rule (static_defender_actions.py:96-128)
                for q, perm_q in rules:
                    if q == port:
                        seen = True
                        patched.append((q, perm))
                    else:
                        patched.append((q, perm_q))
                if not seen:
                    patched.append((port, perm))
                if u_side <= 0.5:
                    nd.fw_in = patched
                else:
                    nd.fw_out = patched
                events += 1
                self.changed_nodes.append(n)
        self.num_events += events                                                          # :419

    def step_attacker_env(self, s, t, vid, kind, u, defender_draws=None):
        """cyberbattle_env.py:299-394 step_attacker_env."""
        if self.done:
            raise RuntimeError("New episode must be started with env.reset()")            # :300-302
        self.stepcount += 1
        if s == t:
            reward, code, recon = self._exploit_local(s, vid, kind, u)
            self.vulnerability_type = "local"
        else:
            reward, code, recon = self._exploit_remote(s, t, vid, kind, u)
            self.vulnerability_type = "remote"
        self.reward = reward
        self.outcome = code
        # update_episode_by_outcome (:397-413)
        if code == C.K_RECON:
            new = 0
            for n in recon:
                if n not in self.discovered_nodes:
                    self.discovered_nodes.append(n)
                    new += 1
            self.discovered_amount += new
        elif code in (C.K_CREDACCESS, C.K_LATERAL):
            self.owned_nodes.append(t)
        elif code in (C.K_COLLECTION, C.K_EXFILTRATION, C.K_DISCOVERY):
            self.discovered_amount += 1
        if self.node_goal and self.interest in self.discovered_nodes and t != self.interest:
            self.reward = 0                                                                # :322-326
        if self.defender:
            self.static_defender_step(defender_draws)                                      # :331-332
        elif self.events_defender:
            self.events_defender_step(defender_draws)
        # end checks (:338-370)
        self.end_episode_reason = 0
        self.truncated = False
        if self.attacker_goal_reached():
            if self.goal == "disruption" or self.cfg.stop_at_goal_reached:
                self.done = True
            self.reward = self.cfg.winning_reward
            self.end_episode_reason = 1
        elif self.check_end_game():
            self.done = True
            self.reward = self.cfg.losing_reward
            self.end_episode_reason = 2
        elif self.proportional_cutoff_coefficient and \
                self.num_iterations >= self.proportional_nodes * self.proportional_cutoff_coefficient:   # :361,:457-460
            self.truncated = True
            self.end_episode_reason = 3
        elif self.num_iterations >= self.episode_iterations:                               # :366
            self.truncated = True
            self.end_episode_reason = 3
        if self.cfg.absolute_reward:
            self.reward = max(0, self.reward)                                              # :379-380
        running = [n for n in self.discovered_nodes if self.nodes[n].status == C.ST_RUNNING]
        self.network_availability = len(running) / len(self.discovered_nodes)              # :389
        self.num_iterations += 1                                                           # :394

    def attacker_goal_reached(self):
        """cyberbattle_env.py:467-514 (network-wide goals)."""
        if self.goal == "control":
            roots = [n for n in self.owned_nodes if self.nodes[n].privilege_level == C.PRIV_ROOT and n != self.starter]
            return len(roots) == self.ownable_count
        if self.goal == "disruption":
            stopped = [n for n in self.discovered_nodes if self.nodes[n].status == C.ST_STOPPED and n != self.starter]
            return len(stopped) == self.disruptable_count
        if self.goal == "discovery":
            disc = [n for n in self.discovered_nodes if n != self.starter]
            data = [n for n in self.discovered_nodes if self.nodes[n].has_data]
            pend = [n for n in self.discovered_nodes if self.nodes[n].data_collected and not self.nodes[n].data_exfiltrated]
            return len(disc) == self.discoverable_count and not data and not pend
        it = self.nodes[self.interest] if self.node_goal else None
        if self.goal == "control_node":                                                    # :488-492
            return self.interest in self.owned_nodes and it.privilege_level == C.PRIV_ROOT
        if self.goal == "discovery_node":                                                  # :493-508
            if self.interest in self.discovered_nodes and it.visible:
                if it.has_data:
                    return bool(it.data_collected and it.data_exfiltrated)
                return True
            return False
        if self.goal == "disruption_node":                                                 # :509-513
            return it.status == C.ST_STOPPED
        return False

    def check_end_game(self):
        """cyberbattle_env.py:438-454: lost when no owned node is Running (or the interest node was killed)."""
        if self.goal in ("control_node", "discovery_node") and self.nodes[self.interest].status == C.ST_STOPPED:
            return True
        return not any(self.nodes[n].status == C.ST_RUNNING for n in self.owned_nodes)

    def get_statistics(self):
        """cyberbattle_env.py:517-524 (14-tuple)."""
        owned = sum(1 for nd in self.nodes if nd.agent_installed)
        disrupted = sum(1 for n in self.discovered_nodes if self.nodes[n].status == C.ST_STOPPED)
        running = sum(1 for n in self.discovered_nodes if self.nodes[n].status == C.ST_RUNNING)
        self.network_availability = running / len(self.discovered_nodes)
        return (owned, len(self.discovered_nodes), self.N - len(self.discovered_nodes), disrupted, self.N,
                self.ownable_count, self.discoverable_count, self.disruptable_count, self.network_availability,
                len(self.overall_reimaged), self.num_events, self.discovered_amount, self.discoverable_amount,
                self.attacker_goal_reached())

    # ---- evolving visible graph + observation (cyberbattle_env_compressed.py) -----------------
    def node_feature_vector(self, n):
        """compressed:319-380 convert_node_info_to_observation + :198-203 (float32, flatten order)."""
        nd, sp = self.nodes[n], self.nodes[n].spec
        M, D = C.MAX_SERVICES, C.VULN_EMB_DIM
        ports = [s.port for s in sp.services]
        fw = [0] * (2 * M)
        if nd.visible:
            for port, perm in nd.fw_in:
                i = ports.index(port) if port in ports else -1
                if i != -1 and i < M:
                    fw[i] = perm
            for port, perm in nd.fw_out:
                i = ports.index(port) if port in ports else -1
                if i != -1 and i < M:
                    fw[M + i] = perm
        running = [0] * M
        fv = [0.0] * D
        if nd.visible:
            acc = np.zeros(D, dtype=np.float64)
            for i, s in enumerate(sp.services):
                if i >= M:
                    break
                running[i] = int(nd.running[i])
                acc = acc + np.asarray(s.fv, dtype=np.float64)
            if len(sp.services) > 0:
                acc = acc / len(sp.services)
            fv = list(acc)
        mean = np.zeros(D, dtype=np.float64)
        if sp.vulns:
            for v in sp.vulns:
                mean = mean + self.vuln_emb[v.vid]
            mean = mean / len(sp.vulns)
        flat = fw + running + [int(nd.visible), int(nd.persistence), int(nd.data_collected), int(nd.data_exfiltrated),
                               int(nd.defense_evasion), int(sp.reimageable), int(nd.privilege_level), int(nd.status),
                               self._value(n), sp.sla_weight] + fv + list(mean)
        return np.array(flat, dtype=np.float32)

    def _add_graph_node(self, n):
        self.graph_nodes.append(n)
        self.node_x[n] = self.node_feature_vector(n)                                       # :206-207

    def _add_edge(self, s, t, vid):
        """compressed:214-246 add_edge_evolving_visible_graph (mean aggregation) incl. the accumulator reset."""
        if t in self.graph_edges.get(s, {}):
            if not self.exploited.get(s).get(t):                                           # :226-227
                self.exploited[s][t] = []
            self.exploited[s][t].append(self.vuln_emb[vid])
        else:
            self.graph_edges.setdefault(s, {})
            self.exploited[s] = {}                                                         # :237 (wipes s's other lists)
            self.exploited[s][t] = [self.vuln_emb[vid]]
        self.graph_edges[s][t] = np.mean(self.exploited[s][t], axis=0)

    def encode(self):
        """compressed:249-306 encode.  *_node goals: the interest node is added to the LIVE graph during an encode, the
        copy being encoded does not have it yet (compressed:254-256), so it takes part from the next encode on."""
        order = list(self.graph_nodes)
        if self.node_goal and self.interest not in self.node_x:
            self._add_graph_node(self.interest)
        pos = {n: i for i, n in enumerate(order)}
        x = torch.from_numpy(np.stack([self.node_x[n] for n in order]))
        src, dst, attrs = [], [], []
        for s in order:                                   # networkx DiGraph.edges order: by source insertion order
            for t, a in self.graph_edges.get(s, {}).items():
                src.append(pos[s])
                dst.append(pos[t])
                attrs.append(a)
        edge_index = torch.tensor([src, dst], dtype=torch.long).view(2, -1)
        if attrs:
            e = torch.from_numpy(np.array(attrs)).float()
        else:
            e = torch.zeros(C.VULN_EMB_DIM, dtype=torch.float32)                           # :258-259 (1-D)
        z = self.gae.forward(x.float(), edge_index, e).numpy()
        self.z_all = z
        running = [n for n in order if self.nodes[n].status == C.ST_RUNNING]
        node_embeddings = {}
        extra = C.NODE_EMB_DIM if self.node_goal else 0
        if not running:                                                                    # :269-274
            return node_embeddings, np.zeros(C.OBS_DIM + extra, dtype=np.float32)
        for n in running:
            node_embeddings[n] = z[pos[n]]
        arr = np.array([node_embeddings[n] for n in node_embeddings], dtype=np.float32)
        obs = np.concatenate([np.average(arr, axis=0), np.max(arr, axis=0), np.min(arr, axis=0)])   # :285-298
        if self.node_goal:                                                                 # :299-305
            if self.interest in running:
                obs = np.concatenate([obs, node_embeddings[self.interest]])
            else:
                obs = np.concatenate([obs, np.zeros(C.NODE_EMB_DIM, dtype=np.float32)])
            if self.interest not in self.discovered_nodes and self.interest in node_embeddings:
                node_embeddings.pop(self.interest)
        return node_embeddings, obs

    def _discrete_features(self):
        return np.array([len(self.discovered_nodes), len(self.owned_nodes)])               # :309-316

    def _reaches(self, targets):
        """Nodes of the evolving visible graph (a DiGraph) from which some node of `targets` can be reached
        (nx.has_path(G, n, target); a node reaches itself)."""
        reach = set(targets)
        changed = True
        while changed:
            changed = False
            for s, outs in self.graph_edges.items():
                if s not in reach and any(t in reach for t in outs):
                    reach.add(s)
                    changed = True
        return reach

    def create_continuous_action_space(self, nodes_to_recalculate=None):
        """compressed:487-523 (+ :526-550).  ``nodes_to_recalculate``
        (precise_action_space_positions, :498-506): pairs whose source or target reaches one of these nodes in the
        visible graph are processed again — their rows are overwritten in place with the current embeddings."""
        run_owned = [n for n in self.owned_nodes if self.nodes[n].status == C.ST_RUNNING]
        run_disc = [n for n in self.discovered_nodes if self.nodes[n].status == C.ST_RUNNING]
        # dict comprehension semantics (:491-494): duplicate keys collapse, first position kept
        run_owned = list(dict.fromkeys(run_owned))
        run_disc = list(dict.fromkeys(run_disc))
        reach = self._reaches(nodes_to_recalculate) if nodes_to_recalculate else ()
        for s in run_owned:
            es = self.node_embeddings[s]
            for t in run_disc:
                if (s, t) in self.processed_pairs and not (s in reach or t in reach):
                    continue
                et = self.node_embeddings[t]
                if s == t:
                    self._add_rows(s, es, t, et, 0)
                self._add_rows(s, es, t, et, 1)
                self.processed_pairs.add((s, t))
        if self.sample_subset_samples:                                                     # :521-522
            self._balance_action_space_by_outcome()

    def _balance_action_space_by_outcome(self):
        """compressed:553-567: group the table's keys by outcome class in first-appearance order; a class with more than
        ``sample_subset_samples`` rows keeps that many (np.random.choice there, philox.subset_keep here: a uniform subset that
        keeps its table order); the table becomes the concatenation of the groups.  Dropped rows are gone for good: their pair
        stays in processed_pairs (only precise_action_space_positions re-adds rows, at the end of the table)."""
        from ccbs_b200.philox import subset_keep, row_identity
        groups = {}
        for i, key in enumerate(self.action_keys):
            groups.setdefault(key[3], []).append(i)
        order = []
        for kind, idx in groups.items():
            if len(idx) > self.sample_subset_samples:
                ks = [self.action_keys[i] for i in idx]
                ident = row_identity([k[0] for k in ks], [k[1] for k in ks], [kind] * len(ks), [self._vloc[k[2]] for k in ks])
                keep = subset_keep(self.philox_seed, self.env_index, self.balance_calls, ident, self.sample_subset_samples)
                self.rows_dropped += len(idx) - len(keep)
                idx = [idx[j] for j in keep]
            order.extend(idx)
        self.action_keys = [self.action_keys[i] for i in order]
        self.action_rows = [self.action_rows[i] for i in order]
        self._row_of_key = {k: i for i, k in enumerate(self.action_keys)}
        self._rows_cache = None
        self.balance_calls += 1

    def _add_rows(self, s, es, t, et, vtype):
        for vid, ri, kind, emb in self.per_node_type[t][vtype]:
            if (s == t and kind == C.K_LATERAL) or kind == C.K_CREDACCESS:                 # :532 (precedence)
                continue
            if self.cfg.remove_all_obstacles and self.goal in ("control", "discovery", "control_node", "discovery_node") \
                    and kind == C.K_DOS:
                continue                                                                   # :536-538
            if self.cfg.remove_main_obstacles and kind == C.K_DOS and t == self.starter:
                continue                                                                   # :541-543
            if self.cfg.remove_main_obstacles and self.node_goal and self.goal != "disruption_node" and \
                    kind == C.K_DOS and t == self.interest:
                continue                                                                   # :545-547
            key, row = (s, t, vid, kind, vtype, ri), np.concatenate((es, et, emb))         # :550
            at = self._row_of_key.get(key)
            if at is None:
                self._row_of_key[key] = len(self.action_keys)
                self.action_keys.append(key)
                self.action_rows.append(row)
            else:
                self.action_rows[at] = row                                                 # "overwrite if changed"
            self._rows_cache = None

    def all_distances(self, action_vector):
        """Distances of the action to every table row, in table order — compressed:571-584: scipy's cosine cdist, or
        np.linalg.norm of (action - rows) with ord 1 / 2 / inf; the action is cast to float32 first (:582)."""
        if self._rows_cache is None:
            self._rows_cache = np.array(self.action_rows)
        seg = np.atleast_2d(np.array(action_vector, dtype=np.float32))
        if self.distance_metric == "cosine":
            return _sp_distance.cdist(seg, self._rows_cache, "cosine").flatten()
        return np.linalg.norm(seg - self._rows_cache, ord={"l1": 1, "l2": 2, "inf": np.inf}[self.distance_metric], axis=1)

    def find_closest_action_embedding(self, action_vector):
        """compressed:570-590."""
        d = self.all_distances(action_vector)
        i = int(np.argmin(d))
        s, t, vid, kind, vtype, ri = self.action_keys[i]
        return s, t, vid, kind, d[i], i

    def step(self, action_vector, uniform, forced=None, defender_draws=None):
        """compressed:389-451.  ``forced`` = (s, t, vid, kind, distance) overrides the decode (used by the
        GPU parity tests to follow a verified near-tie)."""
        if forced is None:
            s, t, vid, kind, dist, row = self.find_closest_action_embedding(action_vector)
        else:
            s, t, vid, kind, dist = forced
            row = -1
        self.last_row = row
        self.step_attacker_env(s, t, vid, kind, float(uniform), defender_draws)
        # update_evolving_visible_graph_after_step (:465-484)
        for n in self.discovered_nodes:
            if n not in self.node_x:
                self._add_graph_node(n)
        if self.outcome in _REFRESH_KINDS:
            self.node_x[t] = self.node_feature_vector(t)
        self.edge_added = self.reward > 0
        if self.edge_added:
            self._add_edge(s, t, vid)
        self.reencoded = kind in _REENCODE_KINDS or self.always_encode                     # :401 (desired outcome / defender)
        if self.reencoded:
            self.node_embeddings, emb = self.encode()
            self.n_encodes += 1
            self.observation = {"graph_embeddings": emb, "discrete_features": self._discrete_features()}
            nodes = None
            if self.precise_positions:                                                     # :419-427
                changes = kind in _REENCODE_KINDS or bool(getattr(self.cfg, "precise_graph_encoding", False))
                nodes = [s, t] if changes else list(self.changed_nodes)                    # defender: around what it changed
            self.create_continuous_action_space(nodes)
        self.reward += self.pen["distance_penalty"] * dist                                 # :430
        info = dict(source_node=s, target_node=t, vulnerability=vid, outcome_kind=kind, outcome_obtained=self.outcome,
                    vulnerability_type=self.vulnerability_type, end_episode_reason=self.end_episode_reason,
                    min_distance_action=dist, step_count=self.stepcount, network_availability=self.network_availability)
        return self.observation, self.reward, self.done or self.truncated, info           # :451

    # ---- state export for bit-exact comparison ---------------------------------------------
    def masks(self):
        """Per-plane bitmask (Python int) in constants.M_* order."""
        m = [0] * C.N_MASKS
        for j, nd in enumerate(self.nodes):
            b = 1 << j
            if nd.agent_installed:
                m[C.M_OWNED] |= b
            if j in self.discovered_nodes:
                m[C.M_DISCOVERED] |= b
            if nd.visible:
                m[C.M_VISIBLE] |= b
            if nd.has_data:
                m[C.M_HAS_DATA] |= b
            if nd.data_collected:
                m[C.M_COLLECTED] |= b
            if nd.data_exfiltrated:
                m[C.M_EXFILTRATED] |= b
            if nd.persistence:
                m[C.M_PERSISTENCE] |= b
            if nd.defense_evasion:
                m[C.M_EVASION] |= b
            if nd.status == C.ST_STOPPED:
                m[C.M_STOPPED] |= b
            if nd.privilege_level >= C.PRIV_USER:
                m[C.M_PRIV_USER] |= b
            if nd.privilege_level == C.PRIV_ROOT:
                m[C.M_PRIV_ROOT] |= b
            if nd.status == C.ST_IMAGING:
                m[C.M_IMAGING] |= b
            if j in self.node_x and int(self.node_x[j][C.F_STATUS]) == C.ST_IMAGING:
                m[C.M_X_IMAGING] |= b
            if self.defender or self.events_defender:        # tracking planes the device only maintains under a defender
                if self._discovered.get(j):
                    m[C.M_EVER_OWNED] |= b
                if j in self._stale:
                    m[C.M_OWN_STALE] |= b
        return m
