"""ctypes binding of libcbsim.so (include/cbsim.h).  There is no fallback: if the shared library is
missing or no CUDA device is usable, construction fails loudly."""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from typing import Optional

import numpy as np

from . import constants as C

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("CBS_LIB", "libcbsim.so"))
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 6

i32, i64, u64, f32, f64 = ct.c_int32, ct.c_int64, ct.c_uint64, ct.c_float, ct.c_double
P = ct.c_void_p


class CbsConfig(ct.Structure):
    _fields_ = [("abi_version", i32), ("device", i32), ("num_envs", i32), ("global_env_offset", i64), ("seed", u64),
                ("goal", i32), ("episode_iterations", i32), ("proportional_cutoff_coefficient", f64),
                ("winning_reward", f64), ("losing_reward", f64), ("absolute_reward", i32), ("stop_at_goal_reached", i32),
                ("remove_main_obstacles", i32), ("remove_all_obstacles", i32), ("switch_interval", i32), ("auto_reset", i32),
                ("rewards", f64 * 10), ("penalties", f64 * 18), ("max_slots", i32), ("max_edges", i32),
                ("decode_margin", f32), ("decode_gemm", i32),
                ("static_defender", i32), ("scan_capacity", i32), ("scan_frequency", i32), ("precise_graph_encoding", i32),
                ("detect_probability", f64), ("precise_action_space_positions", i32), ("distance_metric", i32),
                ("sample_subset_samples", i32), ("random_event_probability", f64)]


_SCENARIO_PTRS = ["sc_num_nodes", "sc_node_off", "sc_port_off", "sc_uvuln_off", "sc_num_uvuln", "sc_instof_off",
                  "sc_discoverable_amount", "sc_init_has_data", "sc_init_visible", "sc_feasible_off", "feasible_starters"]
_SCENARIO_PTRS2 = ["nd_value", "nd_level_at_access", "nd_reimageable", "nd_ownable", "nd_discoverable", "nd_disruptable", "nd_row_off",
                   "outblock", "uvuln_global", "inst_of", "vi_port", "vi_flags", "vi_kinds_any", "vi_kinds_remote",
                   "vi_success", "vi_cost", "vi_recon_any", "vi_recon_remote", "vi_ulocal", "recon_nodes", "row_packed",
                   "row_inst", "vemb32", "vemb64", "vnorm2", "nd_ev_init", "vi_svc_slot", "out_slot"]


class CbsScenarioTables(ct.Structure):
    _fields_ = ([("num_scenarios", i32), ("max_nodes", i32), ("words", i32), ("num_nodes_total", i32), ("num_inst", i32),
                 ("num_rows", i32), ("num_recon", i32), ("num_ports_total", i32), ("num_uvuln_total", i32),
                 ("num_global_vulns", i32), ("num_instof", i64)] + [(n, P) for n in _SCENARIO_PTRS] +
                [("num_feasible", i32), ("sc_interest", P)] + [(n, P) for n in _SCENARIO_PTRS2])


class CbsStateView(ct.Structure):
    """cbs_state_view (include/cbsim.h): every state array of a handle as device pointers"""
    _fields_ = ([(n, i32) for n in ("num_envs", "max_nodes", "words", "mask_pitch", "scalar_pitch", "obs_dim", "slots", "num_masks")] +
                [(n, P) for n in ("masks", "scalars", "disc_order", "owned_order", "pair_slot", "obs", "terminal_obs", "sel", "dist",
                                  "reward64", "last_stats", "stat_accum")])


class CbsReplayLog(ct.Structure):
    """cbs_replay_log (include/cbsim.h)"""
    _fields_ = ([("first_env", i32), ("num_logged", i32)] +
                [(n, P) for n in ("sel", "meta", "reward", "dist", "masks", "disc_order", "owned_order", "counters", "obs", "reset_obs",
                                  "reset_masks", "stats", "force_sel", "force_dist", "force_steps_host")])


_GAE_PTRS = ["node_static", "dyn_proj", "vuln_h", "nn0_b", "bn1_scale", "bn1_shift", "gcn_wt", "bn2_scale", "bn2_shift", "ev_proj"]


class CbsGaeTables(ct.Structure):
    _fields_ = [(n, P) for n in _GAE_PTRS]


# every symbol include/cbsim.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = ["cbs_abi_version", "cbs_create", "cbs_destroy", "cbs_last_error", "cbs_load_scenarios", "cbs_set_scenarios",
           "cbs_set_starter_queue", "cbs_set_action_stride", "cbs_set_actions_prestaged", "cbs_set_defender_draws", "cbs_set_cutoffs", "cbs_reset", "cbs_decode", "cbs_transition",
           "cbs_transition_ksteps", "cbs_observe",
           "cbs_step", "cbs_replay", "cbs_profile_step", "cbs_step_host", "cbs_step_host_async", "cbs_host_sync", "cbs_read_state", "cbs_state_ptr", "cbs_get_state", "cbs_episode_stats", "cbs_reset_stat_accum", "cbs_debug_select_trace", "cbs_debug_observe_trace", "cbs_launch_count",
           "cbs_sync", "cbs_struct_sizes", "cbs_state_bytes", "cbs_capacities"]

# cbs_field
F_MASKS, F_DISC_ORDER, F_OWNED_ORDER, F_SCALARS, F_TERMINAL_OBS, F_OBS, F_LAST_STATS, F_STAT_ACCUM, F_PAIR_SLOT, \
    F_DIST, F_REWARD64, F_ERRFLAG, F_VT, F_OWNED_RAW, F_REIMAGE_LEFT, F_Z_HIST, F_SEL, F_DIVERGENCE, F_EV_CUR, F_EV_X, F_MARGIN_EDGE = range(21)
NUM_SCALARS, NUM_ACCUM = 26, 20
# per-env scalar record (csrc/cbs_types.h enum Scalar): four 32-byte sectors — rewritten every step | list lengths and
# counters | episode constants | misc
(S_FLAGS, S_STEPCOUNT, S_NUM_ITER, S_TOTAL_STEPS, S_OUTCOME, S_SCST, S_EP_RETURN, S_EP_RETURN_HI) = range(8)
(S_N_DISC, S_N_OWNED, S_DISC_AMOUNT, S_N_OWNED_RAW, S_N_REIMAGED, S_N_SLOTS, S_N_EDGES, S_N_ENCODES) = range(8, 16)
(S_SCENARIO, S_STARTER, S_NODE_OFF, S_OWNABLE, S_DISCOVERABLE, S_DISRUPTABLE, S_PROP_NODES,
 S_DISCOVERABLE_AMOUNT) = range(16, 24)
S_EPISODES = 24
S_UVULN_OFF = 25
ACCUM_NAMES = ["episodes", "return_sum", "length_sum", "wins", "lost", "cutoff"] + [f"stat{i}" for i in range(14)]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libcbsim.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "cbsim.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    res = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libcbsim.so failed")
    return LIB_PATH


_lib = None


def load_library():
    """dlopen libcbsim.so and declare the signatures.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the step path)")
    lib = ct.CDLL(LIB_PATH)
    H = P
    lib.cbs_abi_version.restype = ct.c_int
    lib.cbs_create.argtypes = [ct.POINTER(CbsConfig), ct.POINTER(H)]
    lib.cbs_destroy.argtypes = [H]
    lib.cbs_destroy.restype = None
    lib.cbs_last_error.argtypes = [H]
    lib.cbs_last_error.restype = ct.c_char_p
    lib.cbs_load_scenarios.argtypes = [H, ct.POINTER(CbsScenarioTables), ct.POINTER(CbsGaeTables)]
    lib.cbs_set_scenarios.argtypes = [H, P]
    lib.cbs_set_starter_queue.argtypes = [H, P, i32]
    lib.cbs_set_cutoffs.argtypes = [H, i32, f64]
    lib.cbs_set_defender_draws.argtypes = [H, P, P]
    lib.cbs_set_action_stride.argtypes = [H, i32]
    lib.cbs_set_actions_prestaged.argtypes = [H, i32]
    lib.cbs_reset.argtypes = [H, P, P, P]
    lib.cbs_decode.argtypes = [H, P, P, P, P]
    lib.cbs_transition.argtypes = [H, P, P, P, P, P, P, P, P]
    lib.cbs_transition_ksteps.argtypes = [H, P, P, P, i32, P, P, P]
    lib.cbs_observe.argtypes = [H, P, P]
    lib.cbs_step.argtypes = [H, P, P, P, P, P, P, P]
    lib.cbs_replay.argtypes = [H, P, P, i32, ct.POINTER(CbsReplayLog), P]
    lib.cbs_step_host.argtypes = [H, P, P, P, P, P, P]
    lib.cbs_step_host_async.argtypes = [H, P, P, P, P, P, P]
    lib.cbs_host_sync.argtypes = [H]
    lib.cbs_profile_step.argtypes = [H, P, P, P, P]
    lib.cbs_read_state.argtypes = [H, i32, P, i64]
    lib.cbs_read_state.restype = i64
    lib.cbs_get_state.argtypes = [H, P]
    lib.cbs_episode_stats.argtypes = [H, P]
    lib.cbs_state_ptr.argtypes = [H, i32]
    lib.cbs_state_ptr.restype = P
    lib.cbs_reset_stat_accum.argtypes = [H, P]
    lib.cbs_debug_select_trace.argtypes = [H, P]
    lib.cbs_debug_observe_trace.argtypes = [H, P]
    lib.cbs_launch_count.argtypes = [H]
    lib.cbs_launch_count.restype = i64
    lib.cbs_sync.argtypes = [H]
    lib.cbs_struct_sizes.argtypes = [P]
    lib.cbs_state_bytes.argtypes = [H]
    lib.cbs_state_bytes.restype = i64
    lib.cbs_capacities.argtypes = [H, P]
    if lib.cbs_abi_version() != ABI_VERSION:
        raise RuntimeError("libcbsim.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(P)


def make_scenario_struct(tables, goal: int):
    """Fill a CbsScenarioTables from ScenarioTables; returns (struct, keepalive list)."""
    keep = []

    def arr(a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a.ctypes.data_as(P)
    t = CbsScenarioTables()
    t.num_scenarios, t.max_nodes, t.words = tables.num_scenarios, tables.max_nodes, tables.words
    t.num_nodes_total = int(tables.sc_node_off[-1])
    t.num_inst = len(tables.vi_port)
    t.num_rows = len(tables.row_packed)
    t.num_recon = len(tables.recon_nodes)
    t.num_ports_total = int(tables.sc_port_off[-1])
    t.num_uvuln_total = int(tables.sc_uvuln_off[-1])
    t.num_global_vulns = tables.vemb32.shape[0]
    t.num_instof = int(tables.sc_instof_off[-1])
    t.sc_num_nodes = arr(tables.sc_num_nodes, np.int32)
    t.sc_node_off = arr(tables.sc_node_off, np.int32)
    t.sc_port_off = arr(tables.sc_port_off, np.int32)
    t.sc_uvuln_off = arr(tables.sc_uvuln_off, np.int32)
    t.sc_num_uvuln = arr(tables.sc_num_uvuln, np.int32)
    t.sc_instof_off = arr(tables.sc_instof_off, np.int64)
    t.sc_discoverable_amount = arr(tables.sc_discoverable_amount, np.int32)
    t.sc_init_has_data = arr(tables.sc_init_has_data, np.uint32)
    t.sc_init_visible = arr(tables.sc_init_visible, np.uint32)
    t.sc_feasible_off = arr(tables.sc_feasible_off[goal], np.int32)
    t.feasible_starters = arr(tables.feasible_starters[goal], np.int32)
    t.num_feasible = len(tables.feasible_starters[goal])
    t.sc_interest = arr(tables.sc_interest, np.int32) if goal >= C.GOAL_CONTROL_NODE else None
    for name, dt in (("nd_value", np.int32), ("nd_level_at_access", np.uint8), ("nd_reimageable", np.uint8), ("nd_ownable", np.int32),
                     ("nd_discoverable", np.int32), ("nd_disruptable", np.int32), ("nd_row_off", np.int32),
                     ("outblock", np.uint32), ("uvuln_global", np.int32), ("inst_of", np.int32), ("vi_port", np.int32),
                     ("vi_flags", np.uint32), ("vi_kinds_any", np.uint16), ("vi_kinds_remote", np.uint16),
                     ("vi_success", np.float64), ("vi_cost", np.float64), ("vi_recon_any", np.int32),
                     ("vi_recon_remote", np.int32), ("vi_ulocal", np.int32), ("recon_nodes", np.uint8),
                     ("row_packed", np.uint32), ("row_inst", np.int32), ("vemb32", np.float32), ("vemb64", np.float64),
                     ("vnorm2", np.float64), ("nd_ev_init", np.uint16), ("vi_svc_slot", np.uint8), ("out_slot", np.uint8)):
        setattr(t, name, arr(getattr(tables, name), dt))
    return t, keep


def make_gae_struct(gt):
    keep = []
    g = CbsGaeTables()
    for name in _GAE_PTRS:
        a = np.ascontiguousarray(getattr(gt, name), dtype=np.float32)
        keep.append(a)
        setattr(g, name, a.ctypes.data_as(P))
    return g, keep


def make_config(cfg, num_envs: int, device: int = 0, global_env_offset: int = 0, seed: int = 0, auto_reset: bool = True,
                switch_interval: Optional[int] = None, max_slots: int = 0, max_edges: int = 0, decode_margin: float = 0.0,
                decode_gemm: int = 0) -> CbsConfig:
    c = CbsConfig()
    c.abi_version, c.device, c.num_envs = ABI_VERSION, device, num_envs
    c.global_env_offset, c.seed = global_env_offset, seed
    c.goal = C.GOALS[cfg.goal]
    c.episode_iterations = int(cfg.episode_iterations)
    c.proportional_cutoff_coefficient = float(cfg.proportional_cutoff_coefficient or 0)
    c.winning_reward, c.losing_reward = float(cfg.winning_reward), float(cfg.losing_reward)
    c.absolute_reward, c.stop_at_goal_reached = int(cfg.absolute_reward), int(cfg.stop_at_goal_reached)
    c.remove_main_obstacles, c.remove_all_obstacles = int(cfg.remove_main_obstacles), int(cfg.remove_all_obstacles)
    if switch_interval is None:          # not given: the config's (None there = never switch)
        switch_interval = getattr(cfg, "switch_interval", None)
    c.switch_interval, c.auto_reset = (-1 if switch_interval is None else int(switch_interval)), int(auto_reset)
    for i, v in enumerate(cfg.reward_vector()):
        c.rewards[i] = v
    for i, v in enumerate(cfg.penalty_vector()):
        c.penalties[i] = v
    c.max_slots, c.max_edges, c.decode_margin, c.decode_gemm = max_slots, max_edges, decode_margin, decode_gemm
    c.static_defender = {None: 0, "reimage": 1, "events": 2}[getattr(cfg, "static_defender_agent", None)]
    c.random_event_probability = float(getattr(cfg, "random_event_probability", 0.0) or 0.0)
    c.scan_capacity, c.scan_frequency = int(cfg.scan_capacity), int(cfg.scan_frequency)
    c.detect_probability = float(cfg.detect_probability)
    c.precise_graph_encoding = int(bool(cfg.precise_graph_encoding))
    c.precise_action_space_positions = int(bool(getattr(cfg, "precise_action_space_positions", False)))
    c.distance_metric = C.METRICS[getattr(cfg, "distance_metric", "cosine")]
    c.sample_subset_samples = int(getattr(cfg, "sample_subset_samples", 0) or 0)
    return c
