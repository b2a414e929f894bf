"""Environment configuration — same keys as the reference's YAML files
(agents/config/train_config.yaml:3-31, agents/config/rewards_config.yaml) and the keyword
arguments of CyberBattleEnv / CyberBattleCompressedEnv (cyberbattle_env.py:38-60,
cyberbattle_env_compressed.py:74-89)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import constants as C

# agents/config/rewards_config.yaml:4-14 / :72-90 (goal "control")
DEFAULT_REWARDS = {
    "control": dict(value_coefficient=3, cost_coefficient=1, node_discovered_coefficient=25,
                    data_collected_reward=20, data_exfiltrated_reward=20, persistence_reward=20,
                    privilege_escalation_reward=100, acquired_visibility_reward=20, dos_coefficient=-2,
                    defense_evaded_reward=20),
    "disruption": dict(value_coefficient=0.5, cost_coefficient=1, node_discovered_coefficient=10,
                       data_collected_reward=0, data_exfiltrated_reward=0, persistence_reward=0,
                       privilege_escalation_reward=0, acquired_visibility_reward=10, dos_coefficient=5,
                       defense_evaded_reward=10),
    "discovery": dict(value_coefficient=1, cost_coefficient=1, node_discovered_coefficient=100,
                      data_collected_reward=100, data_exfiltrated_reward=100, persistence_reward=0,
                      privilege_escalation_reward=50, acquired_visibility_reward=100, dos_coefficient=-2,
                      defense_evaded_reward=25),
    # agents/config/rewards_config.yaml:15-25,37-47,59-69
    "control_node": dict(value_coefficient=2, cost_coefficient=1, node_discovered_coefficient=15, data_collected_reward=10,
                         data_exfiltrated_reward=10, persistence_reward=10, privilege_escalation_reward=50,
                         acquired_visibility_reward=10, dos_coefficient=-1, defense_evaded_reward=10),
    "disruption_node": dict(value_coefficient=0.5, cost_coefficient=1, node_discovered_coefficient=5, data_collected_reward=0,
                            data_exfiltrated_reward=0, persistence_reward=0, privilege_escalation_reward=0,
                            acquired_visibility_reward=10, dos_coefficient=1, defense_evaded_reward=10),
    "discovery_node": dict(value_coefficient=1, cost_coefficient=1, node_discovered_coefficient=20, data_collected_reward=50,
                           data_exfiltrated_reward=500, persistence_reward=0, privilege_escalation_reward=25,
                           acquired_visibility_reward=50, dos_coefficient=-1, defense_evaded_reward=10),
}
_PEN_COMMON = dict(no_vulnerability_in_node=-10, no_enough_privileges=-10, success_rate_failed=0,
                   no_data_to_collect=-10, no_data_to_exfiltrate=-10, already_persistent=-10,
                   node_already_stopped=-10, node_already_owned=-10, node_already_visible=-10,
                   already_defense_evasion=-10, scanning_unopen_port=-10,
                   privilege_escalation_in_node_not_owned=-10, privilege_escalation_to_level_already_had=-10,
                   outcome_not_valid=-10, blocked_by_local_firewall=-10, blocked_by_remote_firewall=-10,
                   invalid_action=0, distance_penalty=-3)
DEFAULT_PENALTIES = {
    "control": dict(_PEN_COMMON, invalid_action=-50),
    "disruption": dict(_PEN_COMMON),
    "discovery": dict(_PEN_COMMON),
    "control_node": dict(_PEN_COMMON), "disruption_node": dict(_PEN_COMMON), "discovery_node": dict(_PEN_COMMON),
}


@dataclass
class EnvConfig:
    goal: str = "control"
    winning_reward: float = 5000.0
    losing_reward: float = -5000.0
    episode_iterations: int = 200
    proportional_cutoff_coefficient: float = 1
    absolute_reward: bool = False
    stop_at_goal_reached: bool = True
    isolation_filter_threshold: float = 0.1
    remove_main_obstacles: bool = True
    remove_all_obstacles: bool = False
    random_starter_node: bool = True
    # RandomSwitchEnv.switch_interval (cyberbattle_env_switch.py:218-220; 5 in agents/config/train_config.yaml, filled in by
    # from_reference_dicts): a reset draws a new scenario when (episodes finished + 1) % (switch_interval + 1) == 0.
    # None = never switch (every env keeps its scenario); BatchedCyberBattleEnv(switch_interval=...) overrides it.
    switch_interval: Optional[int] = None
    interest_node_value: int = 200        # agents/config/train_config.yaml:16 (value of the node of interest, *_node goals)
    # static defender (_env/static_defender.py): None, "reimage" (ScanAndReimageCompromisedMachines, :27-60) or "events"
    # (ExternalRandomEvents, :63-161: per node and step one of start / stop a service, add / remove a firewall rule, each with
    # random_event_probability).  "events" is restated by the oracle only so far: BatchedCyberBattleEnv refuses it.
    static_defender_agent: Optional[str] = None
    random_event_probability: float = 0.0075   # agents/multi_env/config/train_config.yaml:37-38 (midpoint of [min, max])
    detect_probability: float = 0.05      # train_config.yaml:39-40
    scan_capacity: int = 3                # :41-42
    scan_frequency: int = 3               # :43-44
    # compressed:455-462: when set, every step whose DESIRED outcome is a success class re-encodes the graph - i.e. every step
    precise_graph_encoding: bool = False
    # compressed:86,419-427,498-506: every table-maintaining encode also refreshes the rows of the (source, target) pairs
    # from which the action's source or target node can be reached in the visible graph (their embeddings may have changed)
    precise_action_space_positions: bool = False
    # compressed:82,570-590: metric of the nearest-row decode: 'cosine' (scipy cdist, default), 'l1', 'l2', 'inf' (np.linalg.norm)
    distance_metric: str = "cosine"
    # compressed:83,521-522,553-567: keep at most this many action-table rows per outcome class (0 / False = the whole table).
    # The reference draws the subset with np.random.choice; the oracle and the CUDA path both use the Philox-keyed rule
    # ccbs_b200.philox.subset_keep (pinned against the reference fed the same rule, fixtures s*).
    sample_subset_samples: int = 0
    rewards_dict: Dict[str, float] = field(default_factory=dict)
    penalties_dict: Dict[str, float] = field(default_factory=dict)

    def __post_init__(self):
        self.goal = self.goal.lower()
        if self.goal not in C.GOALS:
            raise ValueError(f"goal '{self.goal}' is not supported by the batched env (supported: {sorted(C.GOALS)})")
        if self.distance_metric not in C.METRICS:
            raise ValueError(f"Unsupported metric '{self.distance_metric}'. Use 'l1', 'l2', 'inf', or 'cosine'.")   # compressed:578-579
        if not self.rewards_dict:
            self.rewards_dict = dict(DEFAULT_REWARDS[self.goal])
        if not self.penalties_dict:
            self.penalties_dict = dict(DEFAULT_PENALTIES[self.goal])
        if self.static_defender_agent not in (None, "reimage", "events"):
            raise ValueError("static_defender_agent must be None, 'reimage' or 'events'")
        if self.static_defender_agent == "events" and self.precise_action_space_positions:
            raise ValueError("precise_action_space_positions cannot be used with the 'events' defender: the reference raises "
                             "networkx.NodeNotFound as soon as an event changes a node that is not in the visible graph "
                             "(nx.has_path on `changed_nodes`, compressed:423-427,498-500)")
        if not self.random_starter_node:
            # cyberbattle_env.py:190-191 takes node 0 and skips the block that computes ownable / discoverable / disruptable
            # counts (:205-217), so the reference's own goal test (:470) raises AttributeError on the first step
            raise ValueError("random_starter_node=False is not supported: the reference never computes the reachable-node "
                             "counts on that branch (cyberbattle_env.py:190-191 vs :205-217) and fails at its first goal test")
        if self.static_defender_agent == "reimage":
            if not (1 <= int(self.scan_capacity) <= C.MAX_SCAN_CAPACITY):
                raise ValueError(f"scan_capacity must be in 1..{C.MAX_SCAN_CAPACITY}")
            if int(self.scan_frequency) < 1:
                raise ValueError("scan_frequency must be >= 1")

    @classmethod
    def from_reference_dicts(cls, train_config: dict, rewards_config: dict, goal: str = "control") -> "EnvConfig":
        """Build from the dicts the reference loads with yaml (agents/train_agent.py:229-239)."""
        keys = {f for f in cls.__dataclass_fields__}
        # reference options the batched env does not implement must not be dropped silently
        train_config = dict(train_config)
        sda = train_config.get("static_defender_agent")
        if sda is not None and not isinstance(sda, (str, bool)):     # already mapped to an object (train_agent.py:393-399)
            sda = {"ScanAndReimageCompromisedMachines": "reimage", "ExternalRandomEvents": "events"}.get(type(sda).__name__, "?")
            obj = train_config["static_defender_agent"]
            for k_obj, k_cfg in (("probability", "detect_probability"), ("scan_capacity", "scan_capacity"),
                                 ("scan_frequency", "scan_frequency")):
                if hasattr(obj, k_obj):
                    train_config[k_cfg] = getattr(obj, k_obj)
        train_config["static_defender_agent"] = sda or None
        if sda and sda not in ("reimage", "events"):
            raise ValueError(f"unknown static defender {sda!r} (the reference has 'reimage' and 'events', _env/static_defender.py)")
        if sda:
            # train_agent.py:395-397 draws the three parameters once per run from [min, max]; the midpoint is used here
            for k in ("detect_probability", "scan_capacity", "scan_frequency"):
                if k not in train_config and f"{k}_min" in train_config:
                    mid = (train_config[f"{k}_min"] + train_config[f"{k}_max"]) / 2
                    train_config[k] = mid if k == "detect_probability" else int(round(mid))
        if train_config.get("pca_components") not in (None, False, 768):
            raise ValueError("only 768-dimensional vulnerability embeddings are implemented")
        kw = {k: v for k, v in train_config.items() if k in keys}
        kw["goal"] = goal
        kw["rewards_dict"] = dict(rewards_config["rewards_dict"][goal])
        kw["penalties_dict"] = dict(rewards_config["penalties_dict"][goal])
        return cls(**kw)

    def reward_vector(self) -> np.ndarray:
        return np.array([float(self.rewards_dict[k]) for k in C.REWARD_KEYS], dtype=np.float64)

    def penalty_vector(self) -> np.ndarray:
        p = dict(self.penalties_dict)
        # the control table spells it `machine_already_stopped` (rewards_config.yaml:79); the code
        # reads `node_already_stopped` (attacker_actions.py:235) on an unreachable branch
        p.setdefault("node_already_stopped", p.get("machine_already_stopped", -10))
        return np.array([float(p[k]) for k in C.PENALTY_KEYS], dtype=np.float64)

    def reference_kwargs(self) -> dict:
        """kwargs for the reference CyberBattleCompressedEnv constructor (used by the golden generator)."""
        pen = dict(self.penalties_dict)
        pen.setdefault("node_already_stopped", pen.get("machine_already_stopped", -10))
        return dict(goal=self.goal, winning_reward=self.winning_reward, losing_reward=self.losing_reward,
                    episode_iterations=self.episode_iterations,
                    proportional_cutoff_coefficient=self.proportional_cutoff_coefficient,
                    absolute_reward=self.absolute_reward, stop_at_goal_reached=self.stop_at_goal_reached,
                    isolation_filter_threshold=self.isolation_filter_threshold,
                    remove_main_obstacles=self.remove_main_obstacles, remove_all_obstacles=self.remove_all_obstacles,
                    random_starter_node=self.random_starter_node, rewards_dict=dict(self.rewards_dict),
                    interest_node_value=self.interest_node_value, switch_interest_node_interval=1,
                    penalties_dict=pen, sample_subset_samples=int(self.sample_subset_samples or 0) or False,
                    static_defender_agent=None,
                    precise_graph_encoding=self.precise_graph_encoding,
                    precise_action_space_positions=self.precise_action_space_positions,
                    distance_metric=self.distance_metric)
