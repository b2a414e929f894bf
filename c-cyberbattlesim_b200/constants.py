"""Shared enumerations and layout constants for the B200 batched continuous env.

Every code here mirrors a class or literal of the reference (cited per item); the CUDA side
(`csrc/cbs_types.h`) carries the same numbers and `tests/test_constants.py` keeps them in sync.
"""

# --- dimensions (reference: agents/config/train_config.yaml:19,32 ; gae/config/train_config.yaml:8,11 ;
#     _env/cyberbattle_env.py:48 ; _env/cyberbattle_env_compressed.py:112-114,365-380)
NODE_EMB_DIM = 64          # GAE output channels
VULN_EMB_DIM = 768         # LM embedding size (pca_components default)
OUTCOME_DIM = 9            # one-hot outcome slots
ACTION_DIM = 2 * NODE_EMB_DIM + VULN_EMB_DIM + OUTCOME_DIM      # 905
MAX_SERVICES = 10          # max_services_per_node
NODE_FEAT_DIM = 3 * MAX_SERVICES + 10 + 2 * VULN_EMB_DIM         # 1576
OBS_DIM = 3 * NODE_EMB_DIM  # mean|max|min readout = 192
NN_CHANNELS = 16           # NNConv edge-network hidden width

# node feature vector offsets (convert_node_info_to_observation, compressed:365-380 + flatten order)
F_FW_IN = 0
F_FW_OUT = 10
F_SVC_RUNNING = 20
F_VISIBLE = 30
F_PERSISTENCE = 31
F_COLLECTED = 32
F_EXFILTRATED = 33
F_EVASION = 34
F_REIMAGEABLE = 35
F_PRIVILEGE = 36
F_STATUS = 37
F_VALUE = 38
F_SLA = 39
F_SVC_FV = 40
F_VULN_MEAN = 40 + VULN_EMB_DIM
DYN_FEATURES = (F_PERSISTENCE, F_COLLECTED, F_EXFILTRATED, F_EVASION, F_PRIVILEGE, F_STATUS)

# --- outcome classes a vulnerability can predict (simulation/model.py:66-193).
# Codes 0..7 equal the LOCAL one-hot index (compressed:595-597); remote one-hot: 0..6 same,
# CredentialAccess -> 7, LateralMove -> 8 (compressed:598-600).
K_DOS = 0
K_DISCOVERY = 1
K_COLLECTION = 2
K_EXFILTRATION = 3
K_RECON = 4
K_EVASION = 5
K_PERSISTENCE = 6
K_PRIVESC = 7
K_CREDACCESS = 8
K_LATERAL = 9
K_EXECUTION = 10
N_KINDS = 11
KIND_NAMES = ["DenialOfService", "Discovery", "Collection", "Exfiltration", "Reconnaissance",
              "DefenseEvasion", "Persistence", "PrivilegeEscalation", "CredentialAccess", "LateralMove",
              "Execution"]
# PredictedResult.outcome_str: the classifier labels the generator maps to outcome classes (utils/encoding_utils.py:129-147,
# simulation/generate_network.py:160-207); printed by RandomSwitchEnv.get_str_info (switch.py:331-333)
KIND_LABELS = ["DOS", "discovery", "collection", "exfiltration", "reconnaissance", "defense evasion", "persistence",
               "privilege escalation", "credential access", "lateral move", "execution"]
# utils/encoding_utils.py:40-62 map_outcome_to_string (info['outcome'])
KIND_INFO_STR = ["DenialOfService", "Discovery", "Collection", "Exfiltration", "Reconnaissance",
                 "DefenseEvasion", "Persistence", "PrivilegeEscalation", "LateralMove-Credential",
                 "LateralMove-Credential", None]

# generate_network.py / encoding_utils.py:129-147 label strings -> kind
LABEL_TO_KIND = {
    "DOS": K_DOS, "discovery": K_DISCOVERY, "collection": K_COLLECTION, "exfiltration": K_EXFILTRATION,
    "reconnaissance": K_RECON, "defense evasion": K_EVASION, "persistence": K_PERSISTENCE,
    "privilege escalation": K_PRIVESC, "credential access": K_CREDACCESS, "lateral move": K_LATERAL,
    "execution": K_EXECUTION,
}


def onehot_index(vtype: int, kind: int):
    """compressed:593-611 map_outcome_to_onehot. vtype 0=local 1=remote. None => row dropped."""
    if kind == K_EXECUTION:
        return None
    if vtype == 0:
        return kind if kind <= K_PRIVESC else None
    if kind <= K_PERSISTENCE:
        return kind
    if kind == K_CREDACCESS:
        return 7
    if kind == K_LATERAL:
        return 8
    return None  # PrivilegeEscalation is not in the remote label list


# --- obtained-outcome codes written by the transition (simulation/attacker_actions.py failure classes)
OC_INVALID_SRC_NOT_OWNED = 16      # InvalidAction, attacker_actions.py:109 / :363
OC_INVALID_TGT_NOT_DISCOVERED = 17  # InvalidAction, :115
OC_SRC_NOT_RUNNING = 18            # NonRunningMachine(0), :121 / :370
OC_TGT_NOT_RUNNING = 19            # NonRunningMachine(1), :127
OC_NO_VULNERABILITY = 20           # NoVulnerability, :133 / :377
OC_NO_PRIVILEGE = 21               # NoEnoughPrivilege, :141 / :386 (and :303 privesc on NoAccess)
OC_OUTCOME_NOT_PRESENT = 22        # OutcomeNonPresent, :153 / :401
OC_PORT_NOT_LISTENING = 23         # NonListeningPort, :161
OC_FW_OUTGOING = 24                # FirewallBlock(1), :170
OC_FW_INCOMING = 25                # FirewallBlock(0), :180
OC_UNSUCCESSFUL = 26               # UnsuccessfulAction, :190 / :409
OC_NO_NEEDED = 27                  # NoNeededAction, :216 / :271
OC_REPEATED = 28                   # RepeatedResult
OC_REMOTE_OUTCOME_LOCAL = 29       # RemoteOutcomeInLocalNode, :345 / :540
OC_NAMES = {
    16: "InvalidAction", 17: "InvalidAction", 18: "NonRunningMachine", 19: "NonRunningMachine",
    20: "NoVulnerability", 21: "NoEnoughPrivilege", 22: "OutcomeNonPresent", 23: "NonListeningPort",
    24: "FirewallBlock", 25: "FirewallBlock", 26: "UnsuccessfulAction", 27: "NoNeededAction",
    28: "RepeatedResult", 29: "RemoteOutcomeInLocalNode",
}

# privilege levels (model.py:60-64) and machine status (model.py:287-291)
PRIV_NONE, PRIV_USER, PRIV_ROOT = 0, 1, 3
ST_STOPPED, ST_RUNNING, ST_IMAGING = 0, 1, 2

# goals (cyberbattle_env.py:467-514): three network-wide goals and their node-specific variants (one "interest node"
# per scenario; the observation then carries that node's embedding as 64 extra floats, compressed:119-125)
GOAL_CONTROL, GOAL_DISCOVERY, GOAL_DISRUPTION = 0, 1, 2
GOAL_CONTROL_NODE, GOAL_DISCOVERY_NODE, GOAL_DISRUPTION_NODE = 3, 4, 5
# decode metrics of find_closest_action_embedding (compressed:571-576); the values are cbs_config.distance_metric
METRIC_COSINE, METRIC_L1, METRIC_L2, METRIC_INF = 0, 1, 2, 3
METRICS = {"cosine": METRIC_COSINE, "l1": METRIC_L1, "l2": METRIC_L2, "inf": METRIC_INF}
GOALS = {"control": GOAL_CONTROL, "discovery": GOAL_DISCOVERY, "disruption": GOAL_DISRUPTION,
         "control_node": GOAL_CONTROL_NODE, "discovery_node": GOAL_DISCOVERY_NODE, "disruption_node": GOAL_DISRUPTION_NODE}


def obs_dim_for_goal(goal: str) -> int:
    """Length of the flat observation: graph_embeddings (192, or 256 for *_node goals) + discrete_features (2)."""
    return OBS_DIM + (NODE_EMB_DIM if goal.endswith("node") else 0) + 2

# end_episode_reason (cyberbattle_env.py:338-370)
END_NONE, END_GOAL, END_LOST, END_CUTOFF = 0, 1, 2, 3

# per-env bitmask planes held in HBM, order shared with csrc/cbs_types.h
M_OWNED = 0        # NodeInfo.agent_installed
M_DISCOVERED = 1   # env.discovered_nodes membership
M_VISIBLE = 2
M_HAS_DATA = 3
M_COLLECTED = 4
M_EXFILTRATED = 5
M_PERSISTENCE = 6
M_EVASION = 7
M_STOPPED = 8      # status == Stopped
M_PRIV_USER = 9    # privilege_level >= LocalUser
M_PRIV_ROOT = 10   # privilege_level == ROOT
# planes that only change under the re-imaging static defender (_env/static_defender.py:27-60)
M_IMAGING = 11      # status == Imaging (static_defender_actions.py:37-52)
M_X_IMAGING = 12    # the node's CACHED feature vector in the evolving visible graph says Imaging (compressed:472-479 refresh rule)
M_EVER_OWNED = 13   # NodeTrackingInformation.last_owned_at is not None (attacker_actions.py:32,87)
M_OWN_STALE = 14    # re-imaged after it was last owned: last_owned_at < last_reimaging (attacker_actions.py:561-573)
N_MASKS = 15
N_MASKS_V1 = 11     # fixtures recorded before the defender planes existed hold the first 11
MASK_NAMES = ["owned", "discovered", "visible", "has_data", "collected", "exfiltrated", "persistence",
              "evasion", "stopped", "priv_user", "priv_root", "imaging", "x_imaging", "ever_owned", "own_stale"]
REIMAGING_DURATION = 15   # static_defender_actions.py:19
MAX_SCAN_CAPACITY = 8     # scan draws per step the device path supports (train_config.yaml:41-42 uses 3)

# reward / penalty vector slots (agents/config/rewards_config.yaml). Order shared with the CUDA side.
REWARD_KEYS = ["value_coefficient", "cost_coefficient", "node_discovered_coefficient",
               "data_collected_reward", "data_exfiltrated_reward", "persistence_reward",
               "privilege_escalation_reward", "acquired_visibility_reward", "dos_coefficient",
               "defense_evaded_reward"]
PENALTY_KEYS = ["no_vulnerability_in_node", "no_enough_privileges", "success_rate_failed",
                "no_data_to_collect", "no_data_to_exfiltrate", "already_persistent",
                "node_already_stopped", "node_already_owned", "node_already_visible",
                "already_defense_evasion", "scanning_unopen_port",
                "privilege_escalation_in_node_not_owned", "privilege_escalation_to_level_already_had",
                "outcome_not_valid", "blocked_by_local_firewall", "blocked_by_remote_firewall",
                "invalid_action", "distance_penalty"]

MAX_NODES = 128  # 4 mask words; config 4 tops out at 100 nodes


def switch_due(episodes_finished: int, switch_interval: int) -> bool:
    """RandomSwitchEnv._check_switch (cyberbattle_env_switch.py:218-220), evaluated by reset() after ``episodes_finished``
    episodes have ended: a new scenario is drawn when ``(episode_count + 1) % (switch_interval + 1) == 0``.  The device applies
    the same test when it resets a finished env in place (k_observe.cu reset_env; ``switch_interval < 0`` = never there)."""
    return (int(episodes_finished) + 1) % (int(switch_interval) + 1) == 0
