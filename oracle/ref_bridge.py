"""TEST INFRASTRUCTURE (build container only) — drive the UNMODIFIED reference env.

Imports ``/root/reference/cyberbattle`` under the stub modules of ``oracle/shims`` and records
traces from the reference's own ``RandomSwitchEnv -> CyberBattleCompressedEnv`` with the
success-rate draws and the starter choice replaced by pre-drawn values.  Nothing here is copied from
the reference; nothing here runs on the GPU box (``/root/reference`` does not exist there).
"""
from __future__ import annotations

import logging
import os
import random as _py_random
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
# the reference tree itself (build container), else the bytecode oracle/build_ref.py compiled from it (travels to the GPU box)
COMPILED_ROOT = os.path.join(_HERE, "_ref", "cyberbattle_ref.zip")
REFERENCE_ROOT = os.environ.get("CBS_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/cyberbattle") else COMPILED_ROOT)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cyberbattle")) or (REFERENCE_ROOT.endswith(".zip") and os.path.isfile(REFERENCE_ROOT))


def reference_kind() -> str:
    """'source' (the read-only tree), 'compiled' (oracle/_ref bytecode of the same files) or 'absent'."""
    if not reference_available():
        return "absent"
    return "compiled" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(COMPILED_ROOT) else "source"


def import_reference():
    """Put the shims and the read-only reference tree on sys.path (never writes bytecode there)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cyberbattle.simulation.model as m                       # noqa: F401
    import cyberbattle.simulation.attacker_actions as aa           # noqa: F401
    import cyberbattle._env.cyberbattle_env as ce                  # noqa: F401
    import cyberbattle._env.cyberbattle_env_compressed as cc       # noqa: F401
    import cyberbattle._env.cyberbattle_env_switch as cs           # noqa: F401
    import cyberbattle.gae.model as gm                             # noqa: F401
    # imported eagerly: its import chain (transformers, ...) draws from the global `random`, which must
    # not happen between random.seed() and Model() in reference_model_from_input_graph
    import cyberbattle.simulation.generate_network                 # noqa: F401
    return dict(model=m, attacker_actions=aa, env=ce, compressed=cc, switch=cs, gae=gm)


def reference_model_from_input_graph(graph: dict, seed: int, **model_kwargs):
    """Run the reference's own scenario generator (model.py:353-420 -> generate_network.py:97) on a
    reference-schema input graph."""
    import networkx as nx
    ref = import_reference()
    G = nx.DiGraph()
    for nid, attrs in graph.items():
        services = []
        for svc in attrs["services"]:
            s = dict(svc)
            s["feature_vector"] = [float(x) for x in svc["feature_vector"]]
            s["vulnerabilities"] = [dict(v, feature_vector=[float(x) for x in v["feature_vector"]])
                                    for v in svc.get("vulnerabilities", [])]
            services.append(s)
        G.add_node(nid, category=attrs["category"], services=services)
    _py_random.seed(seed)
    return ref["model"].Model(network=G, vulnerability_classifier=None, **model_kwargs)


class _FakeRandom:
    """Stands in for the ``random`` module inside attacker_actions / cyberbattle_env."""

    def __init__(self):
        self.next_uniform = None
        self.next_starter = None
        self.uniform_consumed = False
        self.next_choice = None

    def random(self):
        self.uniform_consumed = True
        return float(self.next_uniform)

    def randrange(self, n):
        return int(self.next_starter)

    def choice(self, seq):
        # cyberbattle_env.py:129 switch_interest_node: random.choice(list(nodes)) -> the configured interest node
        return self.next_choice if self.next_choice in seq else seq[0]

    def choices(self, population, k=1):
        # static_defender.py:48 random.choices(list(nodes), k=scan_capacity) -> the pre-drawn node indices
        self.scan_pos = 0
        return [population[int(i)] for i in list(self.next_scan_nodes)[:k]]

    def __getattr__(self, name):
        return getattr(_py_random, name)


def _fake_numpy_random(fake):
    class _R:
        @staticmethod
        def random():
            u = float(fake.next_scan_uniforms[fake.scan_pos])
            fake.scan_pos += 1
            return u
    class _N:
        random = _R()
        def __getattr__(self, name):
            return getattr(np, name)
    return _N()


class _FakeEvents:
    """Stands in for ``random`` and ``numpy`` inside _env/static_defender.py under ExternalRandomEvents (:76-161): per node the
    reference calls random.choice(<4 function names>) (:80), then — unless the node has defense evasion — numpy.random.random()
    for the event test, random.choice(services / ports) and, for firewall events, numpy.random.random() for the side."""
    FUNCTIONS = ["start service", "firewall remove", "stop service", "firewall add"]

    def __init__(self):
        self.draws, self.node, self.calls = None, -1, 0

    def begin_step(self, draws):
        self.draws, self.node, self.calls = np.asarray(draws, np.float64), -1, 0

    def choice(self, seq):
        if list(seq) == self.FUNCTIONS:
            self.node += 1
            self.calls = 0
            return seq[int(self.draws[self.node][0])]
        if len(seq) == 0:
            raise IndexError("Cannot choose from an empty sequence")
        return seq[min(int(float(self.draws[self.node][2]) * len(seq)), len(seq) - 1)]

    def numpy_random(self):
        u = float(self.draws[self.node][1 if self.calls == 0 else 3])
        self.calls += 1
        return u

    def numpy_proxy(self):
        outer = self

        class _R:
            @staticmethod
            def random():
                return outer.numpy_random()

        class _N:
            random = _R()

            def __getattr__(self, name):
                return getattr(np, name)
        return _N()

    def __getattr__(self, name):
        return getattr(_py_random, name)


class ReferenceRunner:
    """One reference ``RandomSwitchEnv(envs_list=[CyberBattleCompressedEnv])`` with controlled randomness."""

    def __init__(self, model, gae_weights, cfg, interest_node=None, subset_vuln_index=None, philox_seed=0, env_index=0):
        import torch
        ref = import_reference()
        self.ref = ref
        logger = logging.getLogger("cbs_ref")
        logger.setLevel(logging.CRITICAL)
        self.fake = _FakeRandom()
        ref["attacker_actions"].random = self.fake
        ref["env"].random = self.fake
        layers = [dict(type="NNConv", NN_channels=16, out_channels=64, activation="ReLU"),
                  dict(type="GCNConv", out_channels=64, activation="ReLU")]
        torch.manual_seed(0)
        enc = ref["gae"].GAEEncoder(1576, layers, 768)
        enc.load_state_dict(gae_weights.state_dict())
        enc.eval()                                                     # agents/train_agent.py:333
        self.fake.next_starter = 0
        ids0 = list(model.network.nodes)
        if interest_node is not None:
            self.fake.next_choice = ids0[int(interest_node)]
            self.fake.next_starter = (int(interest_node) + 1) % len(ids0)   # constructor's reset_env: any starter != interest that passes
        kw = cfg.reference_kwargs()
        if getattr(cfg, "static_defender_agent", None) == "reimage":
            import cyberbattle._env.static_defender as sd
            sd.random = self.fake
            sd.numpy = _fake_numpy_random(self.fake)
            kw["static_defender_agent"] = sd.ScanAndReimageCompromisedMachines(
                cfg.detect_probability, int(cfg.scan_capacity), int(cfg.scan_frequency), logger=logger, verbose=0)
        if getattr(cfg, "static_defender_agent", None) == "events":
            import cyberbattle._env.static_defender as sd
            self.events = _FakeEvents()
            sd.random = self.events
            sd.numpy = self.events.numpy_proxy()
            kw["static_defender_agent"] = sd.ExternalRandomEvents(float(cfg.random_event_probability), logger=logger, verbose=0)
        env = ref["compressed"].CyberBattleCompressedEnv(initial_environment=model, logger=logger, verbose=0, **kw)
        env.set_graph_encoder(enc)
        env.set_pca_components(768)
        self.env = env
        if getattr(cfg, "sample_subset_samples", 0):
            self._install_subset_rule(ref["compressed"], subset_vuln_index, int(philox_seed), int(env_index))
        self.ids = list(model.network.nodes)
        self.index = {n: i for i, n in enumerate(self.ids)}
        self.wrapper = None

    def _install_subset_rule(self, module, vuln_index, seed, env_index):
        """__balance_action_space_by_outcome (compressed:553-567) draws its subset with ``np.random.choice(len(actions), k,
        replace=False)``.  The module-level name ``np`` is replaced by a proxy whose ``random.choice`` applies the documented
        rule (ccbs_b200.philox.subset_keep) to the caller's ``actions`` list; the lifetime balance counter comes from a counting
        wrapper around the (name-mangled) method on the env instance.  Everything else of the method runs unmodified."""
        from ccbs_b200.philox import subset_keep, row_identity
        from ccbs_b200.scenario import _KIND_BY_CLASSNAME
        if vuln_index is None:
            raise ValueError("sample_subset_samples needs subset_vuln_index (oracle.trace.vuln_index(spec))")
        runner, env = self, self.env
        self.balance_calls = 0
        name = "_CyberBattleCompressedEnv__balance_action_space_by_outcome"
        original = getattr(env, name)

        def counted():
            original()
            runner.balance_calls += 1
        setattr(env, name, counted)

        class _Random:
            @staticmethod
            def choice(n, size, replace=False):
                assert replace is False
                actions = sys._getframe(1).f_locals["actions"]
                assert len(actions) == n
                ident = row_identity([runner.index[a[0]] for a in actions], [runner.index[a[1]] for a in actions],
                                     [_KIND_BY_CLASSNAME[type(a[3]).__name__] for a in actions], [vuln_index[a[2]] for a in actions])
                return subset_keep(seed, env_index, runner.balance_calls, ident, size)

            def __getattr__(self, attr):
                return getattr(np.random, attr)

        class _Numpy:
            random = _Random()

            def __getattr__(self, attr):
                return getattr(np, attr)
        module.np = _Numpy()

    def reset(self, starter: int):
        self.fake.next_starter = int(starter)
        if self.wrapper is None:
            np.random.seed(0)
            self.wrapper = self.ref["switch"].RandomSwitchEnv(envs_ids=[0], switch_interval=10 ** 9,
                                                              envs_list=[self.env], verbose=0)
        obs, _ = self.wrapper.reset()
        return obs

    def step(self, action, uniform, defender_draws=None):
        self.fake.next_uniform = float(uniform)
        if defender_draws is not None and len(defender_draws) == 1:        # ExternalRandomEvents: (f, u_event, u_pick, u_side) per node
            self.events.begin_step(defender_draws[0])
        elif defender_draws is not None:
            self.fake.next_scan_nodes, self.fake.next_scan_uniforms = defender_draws
            self.fake.scan_pos = 0
        self.fake.uniform_consumed = False
        obs, reward, done, truncated, info = self.wrapper.step(action)
        return obs, reward, done, truncated, info

    # ---- state extraction --------------------------------------------------------------
    def masks(self):
        import ccbs_b200.constants as C
        env = self.env
        m = [0] * C.N_MASKS
        for j, nid in enumerate(self.ids):
            nd = env.get_node(nid)
            b = 1 << j
            if nd.agent_installed:
                m[C.M_OWNED] |= b
            if nid in env.discovered_nodes:
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
            if nd.status.value == C.ST_STOPPED:
                m[C.M_STOPPED] |= b
            if int(nd.privilege_level) >= C.PRIV_USER:
                m[C.M_PRIV_USER] |= b
            if int(nd.privilege_level) == C.PRIV_ROOT:
                m[C.M_PRIV_ROOT] |= b
            if nd.status.value == C.ST_IMAGING:
                m[C.M_IMAGING] |= b
            g = env.evolving_visible_graph
            if nid in g.nodes and int(g.nodes[nid]["x"][C.F_STATUS]) == C.ST_IMAGING:
                m[C.M_X_IMAGING] |= b
            if env.static_defender_agent:
                track = env._actuator._discovered_nodes.get(nid)
                if track is not None and track.last_owned_at is not None:
                    m[C.M_EVER_OWNED] |= b
                    if nd.last_reimaging is not None and track.last_owned_at < nd.last_reimaging:
                        m[C.M_OWN_STALE] |= b
        return m

    def obtained_code(self):
        """Map env.outcome (a model.VulnerabilityOutcome instance) to constants.OC_* / K_*."""
        import ccbs_b200.constants as C
        from ccbs_b200.scenario import _KIND_BY_CLASSNAME
        o = self.env.outcome
        name = type(o).__name__
        if name in _KIND_BY_CLASSNAME:
            return _KIND_BY_CLASSNAME[name]
        if name == "InvalidAction":
            return C.OC_INVALID_SRC_NOT_OWNED if "source" in o.reason else C.OC_INVALID_TGT_NOT_DISCOVERED
        if name == "NonRunningMachine":
            return C.OC_SRC_NOT_RUNNING if o.source_or_target == 0 else C.OC_TGT_NOT_RUNNING
        if name == "FirewallBlock":
            return C.OC_FW_OUTGOING if o.incoming_or_outgoing == 1 else C.OC_FW_INCOMING
        return {"NoVulnerability": C.OC_NO_VULNERABILITY, "NoEnoughPrivilege": C.OC_NO_PRIVILEGE,
                "OutcomeNonPresent": C.OC_OUTCOME_NOT_PRESENT, "NonListeningPort": C.OC_PORT_NOT_LISTENING,
                "UnsuccessfulAction": C.OC_UNSUCCESSFUL, "NoNeededAction": C.OC_NO_NEEDED,
                "RepeatedResult": C.OC_REPEATED, "RemoteOutcomeInLocalNode": C.OC_REMOTE_OUTCOME_LOCAL}[name]


def make_unpatched_env(model, gae_weights, cfg, seed=0):
    """The reference's own ``RandomSwitchEnv(envs_list=[CyberBattleCompressedEnv])`` exactly as agents/train_agent.py:67 builds
    it — NO randomness is replaced (global ``random`` / ``numpy.random`` seeded like utils/math_utils.py:73-80 set_seeds): the
    thing bench.py times as the CPU reference arm."""
    import torch
    ref = import_reference()
    logger = logging.getLogger("cbs_ref")
    logger.setLevel(logging.CRITICAL)
    _py_random.seed(seed)
    np.random.seed(seed)
    layers = [dict(type="NNConv", NN_channels=16, out_channels=64, activation="ReLU"),
              dict(type="GCNConv", out_channels=64, activation="ReLU")]
    enc = ref["gae"].GAEEncoder(1576, layers, 768)
    enc.load_state_dict(gae_weights.state_dict())
    enc.eval()
    env = ref["compressed"].CyberBattleCompressedEnv(initial_environment=model, logger=logger, verbose=0, **cfg.reference_kwargs())
    env.set_graph_encoder(enc)
    env.set_pca_components(768)
    return ref["switch"].RandomSwitchEnv(envs_ids=[0], switch_interval=10 ** 9, envs_list=[env], verbose=0)


def reference_switch_sequence(models, gae_weights, cfg, switch_interval, picks, n_episodes, seed=0):
    """Scenario index in force during each of the first ``n_episodes`` episodes of the reference's own
    ``RandomSwitchEnv(envs_list=[...], switch_interval=...)`` over several CyberBattleCompressedEnv objects, with
    ``np.random.choice(envs_ids)`` (cyberbattle_env_switch.py:103) answering from ``picks`` in call order.  Also returns how
    many picks were consumed.  Random actions; starters: the first feasible node (rejection loop over randrange)."""
    ref = import_reference()
    runners = [ReferenceRunner(m, gae_weights, cfg) for m in models]
    state = {"i": 0}

    class _Random:
        @staticmethod
        def choice(ids):
            v = ids[int(picks[state["i"]])]
            state["i"] += 1
            return v

        def __getattr__(self, attr):
            return getattr(np.random, attr)

    class _Numpy:
        random = _Random()

        def __getattr__(self, attr):
            return getattr(np, attr)
    module = ref["switch"]
    saved = module.np
    module.np = _Numpy()
    try:
        fake = runners[0].fake                     # the module-level fakes of the LAST runner built serve every env
        for r in runners:
            fake = r.fake
        tries = {"n": 0}
        fake.randrange = lambda n: (tries.__setitem__("n", tries["n"] + 1) or tries["n"]) % n     # walks the nodes until one passes
        wrapper = module.RandomSwitchEnv(envs_ids=list(range(len(models))), switch_interval=switch_interval,
                                         envs_list=[r.env for r in runners], verbose=0)
        rng = np.random.default_rng(seed)
        seq = []
        for _ in range(n_episodes):
            wrapper.reset()
            seq.append(int(wrapper.current_env_index))
            for _ in range(10000):
                fake.next_uniform = float(rng.random())
                _, _, done, truncated, _ = wrapper.step(rng.uniform(-4, 4, 905).astype(np.float32))
                if done or truncated:
                    break
        return seq, state["i"]
    finally:
        module.np = saved
