"""BatchedCyberBattleEnv — B independent continuous envs on one GPU behind libcbsim's C ABI.

Mirrors, per env, ``RandomSwitchEnv(envs_list=[CyberBattleCompressedEnv...])`` of the reference
(_env/cyberbattle_env_switch.py:109-167, _env/cyberbattle_env_compressed.py:158-189,389-451):
``reset()`` -> observations, ``step(actions)`` -> (obs, reward, done, info).  PyTorch supplies device
buffers and the CUDA stream; all env logic runs in the library's kernels.
"""
from __future__ import annotations

import ctypes as ct
from typing import Optional, Sequence

import numpy as np
import torch

from . import constants as C
from . import lib as L
from .config import EnvConfig
from .gae import GaeWeights, fold_gae
from .scenario import ScenarioSpec, ScenarioTables, compile_scenarios


class CbsError(RuntimeError):
    pass


class BatchedCyberBattleEnv:
    def __init__(self, specs: Sequence[ScenarioSpec], gae_weights: GaeWeights, cfg: Optional[EnvConfig] = None,
                 num_envs: int = 1, device: int = 0, scenario_of_env: Optional[np.ndarray] = None, seed: int = 0,
                 global_env_offset: int = 0, auto_reset: bool = True, switch_interval: Optional[int] = None,
                 tables: Optional[ScenarioTables] = None, interest_nodes: Optional[Sequence[int]] = None,
                 gae_tables=None, **cfg_overrides):
        self.cfg = cfg or EnvConfig()
        if getattr(self.cfg, "static_defender_agent", None) == "events":
            from .scenario import check_events_compatible
            for sp in (tables.specs if tables is not None else specs):
                check_events_compatible(sp)
        if not torch.cuda.is_available():
            raise CbsError("BatchedCyberBattleEnv needs a CUDA device (there is no CPU fallback)")
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", device)
        self.lib = L.load_library()
        node_goal = self.cfg.goal.endswith("node")
        if node_goal and tables is None and interest_nodes is None:
            raise CbsError("*_node goals need interest_nodes (one node index per scenario; the reference draws it with "
                           "random.choice when the env object is built, cyberbattle_env.py:127-131)")
        self.tables = tables if tables is not None else compile_scenarios(
            specs, self.cfg.isolation_filter_threshold, interest_nodes=interest_nodes if node_goal else None,
            interest_node_value=self.cfg.interest_node_value if node_goal else None)
        self.obs_dim = C.obs_dim_for_goal(self.cfg.goal)
        self.gae_tables = gae_tables if gae_tables is not None else fold_gae(self.tables, gae_weights)
        ccfg = L.make_config(self.cfg, self.num_envs, device=device, global_env_offset=global_env_offset, seed=seed,
                             auto_reset=auto_reset, switch_interval=switch_interval, **cfg_overrides)
        self._h = ct.c_void_p()
        rc = self.lib.cbs_create(ct.byref(ccfg), ct.byref(self._h))
        if rc != 0:
            raise CbsError(f"cbs_create failed ({rc}): {self.lib.cbs_last_error(None).decode()}")
        st, keep1 = L.make_scenario_struct(self.tables, C.GOALS[self.cfg.goal])
        gt, keep2 = L.make_gae_struct(self.gae_tables)
        self._check(self.lib.cbs_load_scenarios(self._h, ct.byref(st), ct.byref(gt)))
        del keep1, keep2
        if scenario_of_env is None:
            scenario_of_env = np.arange(self.num_envs, dtype=np.int32) % self.tables.num_scenarios
        self.scenario_of_env = np.ascontiguousarray(scenario_of_env, dtype=np.int32)
        self._check(self.lib.cbs_set_scenarios(self._h, self.scenario_of_env.ctypes.data_as(ct.c_void_p)))
        caps = (ct.c_int32 * 8)()
        self.lib.cbs_capacities(self._h, caps)
        self.ncap, self.slots, self.ecap, self.tensor_core_decode = caps[0], caps[1], caps[2], bool(caps[3])
        self.vt_stride = caps[4]
        self._mask_pitch, self._scalar_pitch = caps[6], caps[7]
        assert caps[5] == self.obs_dim
        B = self.num_envs
        self._act_stride = C.ACTION_DIM
        with torch.cuda.device(self.device):
            # zero-copy view of the library's observation cache [B, 194]
            self.obs = _tensor_from_ptr(self.lib.cbs_state_ptr(self._h, L.F_OBS), (B, self.obs_dim), torch.float32,
                                        self.device, self)
            self.reward = torch.zeros(B, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(B, dtype=torch.uint8, device=self.device)
            self.truncated = torch.zeros(B, dtype=torch.uint8, device=self.device)
            self.outcome = torch.zeros(B, dtype=torch.uint8, device=self.device)
            self.info = torch.zeros(B, 8, dtype=torch.int32, device=self.device)
            # zero-copy views of the library's decode result: decode() writes them, transition(self.sel, self.dist)
            # reads them in place (no copy of the selection on either side)
            self.sel = _tensor_from_ptr(self.lib.cbs_state_ptr(self._h, L.F_SEL), (B, 4), torch.int32, self.device, self)
            self.dist = _tensor_from_ptr(self.lib.cbs_state_ptr(self._h, L.F_DIST), (B,), torch.float64, self.device, self)

    # ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise CbsError(f"libcbsim error {rc}: {self.lib.cbs_last_error(self._h).decode()}")

    def _stream(self):
        return ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return None if t is None else ct.c_void_p(t.data_ptr())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cbs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-facing surface ------------------------------------------------------------
    def set_starter_queue(self, queue: Optional[np.ndarray]):
        """queue[num_envs, qlen] explicit starter nodes per episode (tests); None = random feasible starters."""
        if queue is None:
            self._check(self.lib.cbs_set_starter_queue(self._h, None, 0))
        else:
            q = np.ascontiguousarray(queue, dtype=np.int32)
            assert q.shape[0] == self.num_envs
            self._check(self.lib.cbs_set_starter_queue(self._h, q.ctypes.data_as(ct.c_void_p), q.shape[1]))

    def set_defender_draws(self, scan_nodes: Optional[torch.Tensor], detect_uniforms: Optional[torch.Tensor]):
        """Test hook: replace the static defender's random draws (static_defender.py:48,53) by device tensors
        scan_nodes int32[B, scan_capacity] / detect_uniforms float32[B, scan_capacity]; the tensors are read by every
        following step (update them in place), ``None, None`` restores the Philox streams."""
        if self.cfg.static_defender_agent == "events" and detect_uniforms is not None:
            # ExternalRandomEvents: detect_uniforms = float32[B, max_nodes, 4] per node (function index, event / pick / side uniform)
            assert detect_uniforms.shape == (self.num_envs, self.ncap, 4) and detect_uniforms.dtype == torch.float32
            assert detect_uniforms.is_contiguous() and detect_uniforms.device == self.device
            self._def_draws = (detect_uniforms,)
            self._check(self.lib.cbs_set_defender_draws(self._h, None, self._p(detect_uniforms)))
            return
        if scan_nodes is None and detect_uniforms is None:
            self._def_draws = None
            self._check(self.lib.cbs_set_defender_draws(self._h, None, None))
            return
        k = int(self.cfg.scan_capacity)
        assert scan_nodes.shape == (self.num_envs, k) and detect_uniforms.shape == (self.num_envs, k)
        assert scan_nodes.dtype == torch.int32 and detect_uniforms.dtype == torch.float32
        assert scan_nodes.is_contiguous() and detect_uniforms.is_contiguous() and scan_nodes.device == self.device
        self._def_draws = (scan_nodes, detect_uniforms)          # keep them alive
        self._check(self.lib.cbs_set_defender_draws(self._h, self._p(scan_nodes), self._p(detect_uniforms)))

    def set_actions_prestaged(self, on: bool = True):
        """Declare that the action tensors passed to step() / decode() are complete long before the call (a pre-generated ring,
        a recorded trace) - not produced by a kernel enqueued just before it.  Lets the contraction of the next step read them
        while this step's observe kernel drains (include/cbsim.h: cbs_set_actions_prestaged).  Off by default."""
        self._check(self.lib.cbs_set_actions_prestaged(self._h, 1 if on else 0))

    def set_cut_off(self, cut_off: int):
        """cyberbattle_env_switch.py:198-199"""
        self.cfg.episode_iterations = int(cut_off)
        self._check(self.lib.cbs_set_cutoffs(self._h, int(cut_off), float(self.cfg.proportional_cutoff_coefficient or 0)))

    def set_proportional_cutoff_coefficient(self, coefficient: float):
        """cyberbattle_env_switch.py:202-203"""
        self.cfg.proportional_cutoff_coefficient = coefficient
        self._check(self.lib.cbs_set_cutoffs(self._h, int(self.cfg.episode_iterations), float(coefficient or 0)))

    def reset(self, env_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all envs (or those with env_mask != 0).  Returns obs[B, 194] (device tensor, reused)."""
        if env_mask is not None:
            env_mask = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self._check(self.lib.cbs_reset(self._h, self._p(env_mask), None, self._stream()))
        return self.obs

    def decode(self, actions: torch.Tensor):
        actions = self._actions(actions)
        self._check(self.lib.cbs_decode(self._h, self._p(actions), None, None, self._stream()))
        return self.sel, self.dist

    def transition(self, sel: torch.Tensor, dist: Optional[torch.Tensor] = None, uniforms: Optional[torch.Tensor] = None):
        if sel is not self.sel:
            sel = sel.to(device=self.device, dtype=torch.int32).contiguous()
        if dist is not None:
            dist = dist.to(device=self.device, dtype=torch.float64).contiguous()
        uniforms = self._uniforms(uniforms)
        self._check(self.lib.cbs_transition(self._h, self._p(sel), self._p(dist), self._p(uniforms), self._p(self.reward),
                                            self._p(self.done), self._p(self.truncated), self._p(self.outcome), self._stream()))
        return self.reward, self.done, self.truncated, self.outcome

    def transition_ksteps(self, sel: torch.Tensor, dist: Optional[torch.Tensor] = None, uniforms: Optional[torch.Tensor] = None):
        """``cbs_transition_ksteps``: K pre-decoded actions per env in one launch (sel int32[K, B, 4], dist float64[K, B],
        uniforms float32[K, B] or None for Philox).  Returns (reward float32[K, B], done uint8[K, B]).  The visible graph / action
        table are not advanced (no observe between the steps): reset() before stepping normally again."""
        K = int(sel.shape[0])
        assert sel.shape == (K, self.num_envs, 4) and sel.dtype == torch.int32 and sel.is_contiguous() and sel.device == self.device
        if dist is not None:
            assert dist.shape == (K, self.num_envs) and dist.dtype == torch.float64 and dist.is_contiguous()
        if uniforms is not None:
            assert uniforms.shape == (K, self.num_envs) and uniforms.dtype == torch.float32 and uniforms.is_contiguous()
        reward = torch.empty(K, self.num_envs, dtype=torch.float32, device=self.device)
        done = torch.empty(K, self.num_envs, dtype=torch.uint8, device=self.device)
        self._check(self.lib.cbs_transition_ksteps(self._h, self._p(sel), self._p(dist), self._p(uniforms), K, self._p(reward),
                                                   self._p(done), self._stream()))
        return reward, done

    def observe(self) -> torch.Tensor:
        self._check(self.lib.cbs_observe(self._h, None, self._stream()))
        return self.obs

    def step(self, actions: torch.Tensor, uniforms: Optional[torch.Tensor] = None, want_info: bool = True):
        """One batched step with device tensors.  Returns (obs[B,194], reward[B], done[B] (= done|truncated,
        compressed:451), info[B,8] int32).  Finished envs are reset in place when auto_reset is on; their last
        observation is available from :meth:`terminal_obs`."""
        actions = self._actions(actions)
        uniforms = self._uniforms(uniforms)
        self._check(self.lib.cbs_step(self._h, self._p(actions), self._p(uniforms), None, self._p(self.reward),
                                      self._p(self.done), self._p(self.info) if want_info else None, self._stream()))
        return self.obs, self.reward, self.done, self.info

    def replay(self, actions: torch.Tensor, uniforms: Optional[torch.Tensor] = None, log_envs: Optional[slice] = None,
               forced: Optional[dict] = None) -> dict:
        """``cbs_replay``: T whole steps over pre-staged inputs (actions float32[T, B, 905], uniforms float32[T, B] or None for
        Philox) without a host round trip per step; every step's decoded action, outcome, reward, distance, state records and
        observation of the envs in ``log_envs`` (default: all) go to a device log, returned as a dict of numpy arrays
        [T, n_logged, ...] after ONE synchronisation.  Keys: sel, code, done, truncated, reason, step_count, episode, reward,
        dist, masks (uint32 [T, n, N_MASKS, words]), disc_order, owned_order, counters, n_disc, n_owned, scenario, obs, reset_obs,
        reset_masks, stats.  ``forced`` = {step: (sel[4], distance)}: at those steps EVERY env takes that decoded action instead
        of its own (parity tests following a verified near-tie)."""
        T = int(actions.shape[0])
        assert actions.shape == (T, self.num_envs, C.ACTION_DIM) and actions.dtype == torch.float32 and actions.is_contiguous()
        assert actions.device == self.device
        if self._act_stride != C.ACTION_DIM:
            self._act_stride = C.ACTION_DIM
            self._check(self.lib.cbs_set_action_stride(self._h, self._act_stride))
        if uniforms is not None:
            assert uniforms.shape == (T, self.num_envs) and uniforms.dtype == torch.float32 and uniforms.is_contiguous()
        lo, hi, _ = (log_envs or slice(None)).indices(self.num_envs)
        n = hi - lo
        defender = bool(self.cfg.static_defender_agent)
        olen = (2 if defender else 1) * self.ncap
        z = lambda *shape, dtype: torch.zeros(T, n, *shape, dtype=dtype, device=self.device)   # noqa: E731
        bufs = dict(sel=z(4, dtype=torch.int32), meta=z(4, dtype=torch.int32), reward=z(dtype=torch.float64), dist=z(dtype=torch.float64),
                    masks=z(self._mask_pitch, dtype=torch.int32), disc_order=z(self.ncap, dtype=torch.uint8),
                    owned_order=z(olen, dtype=torch.uint8), counters=z(8, dtype=torch.int32), obs=z(self.obs_dim, dtype=torch.float32),
                    reset_obs=z(self.obs_dim, dtype=torch.float32), reset_masks=z(self._mask_pitch, dtype=torch.int32),
                    stats=z(14, dtype=torch.float64))
        log = L.CbsReplayLog()
        log.first_env, log.num_logged = lo, n
        for k, v in bufs.items():
            setattr(log, k, v.data_ptr())
        if forced:
            f_sel = torch.full((T, self.num_envs, 4), -1, dtype=torch.int32)
            f_dist = torch.zeros(T, self.num_envs, dtype=torch.float64)
            steps = np.zeros(T, dtype=np.uint8)
            for t, (sel4, d) in forced.items():
                f_sel[t] = torch.as_tensor(np.asarray(sel4, np.int32))
                f_dist[t] = float(d)
                steps[t] = 1
            f_sel, f_dist = f_sel.to(self.device), f_dist.to(self.device)
            log.force_sel, log.force_dist, log.force_steps_host = f_sel.data_ptr(), f_dist.data_ptr(), steps.ctypes.data
        self._check(self.lib.cbs_replay(self._h, self._p(actions), self._p(uniforms), T, ct.byref(log), self._stream()))
        self.sync()
        out = {k: v.cpu().numpy() for k, v in bufs.items()}
        w = self.tables.words
        meta, cnt = out.pop("meta"), out["counters"]
        unpack = lambda m: np.ascontiguousarray(m.view(np.uint32)[..., :C.N_MASKS * w].reshape(T, n, C.N_MASKS, w))   # noqa: E731
        out["masks"], out["reset_masks"] = unpack(out["masks"]), unpack(out["reset_masks"])
        out["code"], flags, out["step_count"], out["episode"] = meta[..., 0], meta[..., 1], meta[..., 2], meta[..., 3]
        out["done"], out["truncated"], out["reason"], out["scenario"] = flags & 1, (flags >> 1) & 1, (flags >> 2) & 3, flags >> 8
        out["n_disc"], out["n_owned"] = cnt[..., 7] & 0xFFFF, cnt[..., 7] >> 16
        out["counters"] = cnt[..., :7]
        return out

    def profile_step(self, actions: torch.Tensor) -> dict:
        """One step with CUDA events between the launches; returns milliseconds per kernel group."""
        actions = self._actions(actions)
        out = (ct.c_float * 5)()
        self._check(self.lib.cbs_profile_step(self._h, self._p(actions), None, out, self._stream()))
        return {"decode_gemm": out[0], "decode_select_transition": out[1], "observe": out[2]}

    def step_host(self, actions: np.ndarray, uniforms: Optional[np.ndarray], obs: np.ndarray, reward: np.ndarray,
                  done: np.ndarray, info: Optional[np.ndarray] = None):
        """The same step through HOST buffers (numpy, ideally pinned): copies in, steps, copies out, synchronises."""
        assert actions.dtype == np.float32 and actions.shape == (self.num_envs, C.ACTION_DIM) and actions.flags.c_contiguous
        v = lambda a: None if a is None else a.ctypes.data_as(ct.c_void_p)  # noqa: E731
        self._check(self.lib.cbs_step_host(self._h, v(actions), v(uniforms), v(obs), v(reward), v(done), v(info)))

    def step_host_async(self, actions: np.ndarray, uniforms: Optional[np.ndarray], obs: np.ndarray, reward: np.ndarray,
                        done: np.ndarray, info: Optional[np.ndarray] = None):
        """:meth:`step_host` without the final wait (``cbs_step_host_async``); pair with :meth:`host_sync`.  The buffers
        must stay alive, unmodified, until then."""
        assert actions.dtype == np.float32 and actions.shape == (self.num_envs, C.ACTION_DIM) and actions.flags.c_contiguous
        v = lambda a: None if a is None else a.ctypes.data_as(ct.c_void_p)  # noqa: E731
        self._check(self.lib.cbs_step_host_async(self._h, v(actions), v(uniforms), v(obs), v(reward), v(done), v(info)))

    def host_sync(self):
        self._check(self.lib.cbs_host_sync(self._h))

    def _actions(self, actions):
        """Accepts a dense [B, 905] tensor or a [B, 905] view of a wider row-pitched buffer (stride(0) >= 905,
        stride(1) == 1); the pitch is forwarded so that a 16-byte-multiple pitch is read in place by TMA."""
        if actions.shape != (self.num_envs, C.ACTION_DIM):
            raise ValueError(f"actions must have shape ({self.num_envs}, {C.ACTION_DIM})")
        actions = actions.to(device=self.device, dtype=torch.float32)
        if actions.stride(1) != 1 or actions.stride(0) < C.ACTION_DIM:
            actions = actions.contiguous()
        if actions.stride(0) != self._act_stride:
            self._act_stride = int(actions.stride(0))
            self._check(self.lib.cbs_set_action_stride(self._h, self._act_stride))
        return actions

    def _uniforms(self, uniforms):
        if uniforms is None:
            return None
        return uniforms.to(device=self.device, dtype=torch.float32).contiguous()

    # ---- state access ---------------------------------------------------------------------
    def sync(self):
        self._check(self.lib.cbs_sync(self._h))

    def read(self, field: int, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        n = self.lib.cbs_read_state(self._h, field, out.ctypes.data_as(ct.c_void_p), out.nbytes)
        if n < 0:
            self._check(int(n))
        return out

    def masks(self) -> np.ndarray:
        """uint32[N_MASKS, words, B]"""
        raw = self.read(L.F_MASKS, np.uint32, (self.num_envs, self._mask_pitch))       # env-major records on the device
        w = self.tables.words
        return np.ascontiguousarray(raw[:, :C.N_MASKS * w].reshape(self.num_envs, C.N_MASKS, w).transpose(1, 2, 0))

    def scalars(self) -> np.ndarray:
        """int32[NUM_SCALARS, B], indexed by lib.S_*"""
        raw = self.read(L.F_SCALARS, np.int32, (self._scalar_pitch // 8, self.num_envs, 8))   # sector-major on the device
        return np.ascontiguousarray(raw.transpose(0, 2, 1).reshape(self._scalar_pitch, self.num_envs)[:L.NUM_SCALARS])

    def disc_order(self) -> np.ndarray:
        return self.read(L.F_DISC_ORDER, np.uint8, (self.num_envs, self.ncap))

    def owned_order(self) -> np.ndarray:
        return self.read(L.F_OWNED_ORDER, np.uint8, (self.num_envs, self.ncap))

    def owned_raw(self) -> np.ndarray:
        """env.owned_nodes exactly as the reference keeps it under a static defender (removals, duplicates): uint8[B, 2*ncap],
        length in scalars()[S_N_OWNED_RAW]."""
        if not self.cfg.static_defender_agent:
            raise RuntimeError("owned_raw() is only maintained with a static defender; use owned_order()")
        return self.read(L.F_OWNED_RAW, np.uint8, (self.num_envs, 2 * self.ncap))

    def terminal_obs(self) -> np.ndarray:
        return self.read(L.F_TERMINAL_OBS, np.float32, (self.num_envs, self.obs_dim))

    def last_stats(self) -> np.ndarray:
        """get_statistics() 14-tuple of the last finished episode of every env (cyberbattle_env.py:517-524)."""
        return self.read(L.F_LAST_STATS, np.float64, (self.num_envs, 14))

    def vt(self) -> np.ndarray:
        """float32[B, vt_stride]: a_v . v_u of the last decode (columns >= global_vulns are padding)."""
        return self.read(L.F_VT, np.float32, (self.num_envs, self.vt_stride))

    def reward64(self) -> np.ndarray:
        return self.read(L.F_REWARD64, np.float64, (self.num_envs,))

    def distances(self) -> np.ndarray:
        """float64[B]: distance of the last decoded action to its table row (info['min_distance_action'], compressed:449)."""
        return self.read(L.F_DIST, np.float64, (self.num_envs,))

    def events_state(self, cached: bool = False) -> np.ndarray:
        """Events defender: uint16[B, max_nodes, 4] = per node the { running services, incoming BLOCK, outgoing BLOCK, - } bit sets
        over the node's service slots; ``cached`` = as held by the node's feature vector in the visible graph."""
        return self.read(L.F_EV_X if cached else L.F_EV_CUR, np.uint16, (self.num_envs, self.ncap, 4))

    def margin_edge_count(self) -> int:
        """Decodes so far whose float64 winner sat in the outer half of the float32 re-score margin (see CBS_F_MARGIN_EDGE)."""
        return int(self.read(L.F_MARGIN_EDGE, np.int32, (1,))[0])

    def divergence_count(self) -> int:
        """Env-steps so far at which the reference itself would have raised (see CBS_F_DIVERGENCE in include/cbsim.h)."""
        return int(self.read(L.F_DIVERGENCE, np.int32, (1,))[0])

    def stat_accum(self) -> dict:
        a = self.read(L.F_STAT_ACCUM, np.float64, (L.NUM_ACCUM,))
        return dict(zip(L.ACCUM_NAMES, a.tolist()))

    def stat_accum_tensor(self) -> torch.Tensor:
        """Zero-copy float64[20] view of the device-side episode accumulators (input of the NCCL all-reduce)."""
        ptr = self.lib.cbs_state_ptr(self._h, L.F_STAT_ACCUM)
        return _tensor_from_ptr(ptr, (L.NUM_ACCUM,), torch.float64, self.device, self)

    def reset_stat_accum(self):
        self._check(self.lib.cbs_reset_stat_accum(self._h, self._stream()))

    @property
    def launch_count(self) -> int:
        return int(self.lib.cbs_launch_count(self._h))

    @property
    def state_bytes(self) -> int:
        return int(self.lib.cbs_state_bytes(self._h))


def _tensor_from_ptr(ptr, shape, dtype, device, owner):
    """Wrap raw device memory as a torch tensor through the CUDA array interface."""
    itemsize = torch.tensor([], dtype=dtype).element_size()
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]

    class _Holder:
        pass
    hld = _Holder()
    hld.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2,
                                        strides=None)
    hld._owner = owner
    t = torch.as_tensor(hld, device=device)
    assert t.element_size() == itemsize
    return t
