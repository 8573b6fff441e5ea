"""BatchedEngine: thin Python owner of one libev2b handle (one per GPU).

PyTorch is plumbing here: it allocates the output buffers, provides the CUDA stream, and wraps
the engine's struct-of-arrays state as zero-copy `torch.Tensor` views for RL code.  All compute
is in the hand-written CUDA kernels behind the C ABI (include/ev2b.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .scenario import LUT_LEN, SESSION_F64_FIELDS, SESSION_INT_FIELDS, Scenario, Topology

REWARD_KINDS = {None: 0, "none": 0, "SquaredTrackingErrorReward": 1,
                "ProfitMax_TrPenalty_UserIncentives": 2, "profit_maximization": 3, "V2G_grid_full_reward": 4,
                "V2G_grid_simple_reward": 5, "SqTrError_TrPenalty_UserIncentives": 6, "SimpleReward": 7,
                "MinimizeTrackerSurplusWithChargeRewards": 8, "V2G_profitmax": 9, "V2G_costs_simple": 10,
                "V2G_profitmaxV2": 11, "Grid_V2G_profitmaxV2": 12, "pst_V2G_profitmaxV2": 13,
                "SquaredTrackingErrorRewardWithPenalty": 14}
STATE_KINDS = {None: 0, "none": 0, "PublicPST": 1, "V2G_profit_max": 2, "V2G_profit_max_loads": 3, "V2G_grid_state": 4}

ST_DONE, ST_AMPS_OVERFLOW, ST_WAS_DONE = 1, 2, 4
KPI_NAMES = ("total_reward", "total_profits", "total_energy_charged", "total_energy_discharged",
             "total_transformer_overload", "total_ev_served", "sat_sum", "tracking_error",
             "energy_tracking_error_steps", "power_tracker_violation", "total_evs_spawned", "invalid_actions",
             "steps")

STAT_NAMES = ("total_ev_served", "total_profits", "total_energy_charged", "total_energy_discharged",
              "average_user_satisfaction", "power_tracker_violation", "tracking_error", "energy_tracking_error",
              "energy_user_satisfaction", "std_energy_user_satisfaction", "min_energy_user_satisfaction",
              "total_steps_min_emergency_battery_capacity_violation", "total_transformer_overload",
              "battery_degradation", "battery_degradation_calendar", "battery_degradation_cycling", "total_reward")

_OUT_SPECS = {  # name -> (dtype, per-env shape key)
    "reward": ("float64", ()), "status": ("int32", ()), "obs": ("float32", ("D",)),
    "cs_power": ("float32", ("C",)), "cs_current": ("float32", ("C",)),
    "tr_power": ("float64", ("Tr",)), "tr_overload": ("float64", ("Tr",)), "total_costs": ("float64", ()),
    "action_mask": ("uint8", ("P",)), "dep_sat": ("float64", ("P",)), "dep_cap": ("float64", ("P",)),
    "port_energy": ("float32", ("P",)), "node_voltage": ("float64", ("N",)),
    # per-episode histories, time-major (include/ev2b.h); BatchedEngine.histories() gives the reference's [.,T] shapes
    "hist_cs_power": ("float32", ("T", "C")), "hist_cs_current": ("float32", ("T", "C")),
    "hist_tr_overload": ("float64", ("T", "Tr")), "hist_usage": ("float64", ("T",)),
}


class EngineError(RuntimeError):
    pass


def _fn_name(f) -> Optional[str]:
    if f is None or isinstance(f, str):
        return f
    return getattr(f, "__name__", None)


def topology_view(topo: Topology):
    """`ev2b_topology` (include/ev2b.h) over the arrays of a Topology; returns (view, arrays to keep alive)."""
    keep = [topo.cs_n_ports, topo.cs_tr, topo.cs_phases, topo.cs_imax, topo.cs_imin, topo.cs_imax_dis,
            topo.cs_imin_dis, topo.cs_voltage]
    tv = _lib.TopologyView(*[a.ctypes.data_as(t) for a, (_, t) in zip(keep, _lib.TopologyView._fields_)])
    if topo.n_bus:
        gk = np.ascontiguousarray(topo.grid_K).view(np.float64).reshape(-1)
        gl = np.ascontiguousarray(topo.grid_L).view(np.float64).reshape(-1)
        keep += [gk, gl]
        tv.n_bus, tv.grid_s_base = topo.n_bus, float(topo.grid_s_base)
        tv.grid_K, tv.grid_L = gk.ctypes.data_as(_lib._pd), gl.ctypes.data_as(_lib._pd)
    return tv, keep


def scenarios_view(topo: Topology, scenarios: Sequence[Scenario]):
    """`ev2b_scenarios` (include/ev2b.h) over a bank of Scenario objects; returns (view, arrays to keep alive)."""
    n = len(scenarios)
    T, Tr = topo.T, topo.Tr
    for sc in scenarios:
        sc.normalise()
        if sc.charge_price.shape != (T,) or sc.tr_infl.shape != (Tr, T):
            raise EngineError("scenario shape does not match the engine's topology")
    v = _lib.ScenariosView()
    keep = []

    def put(name, arr, ptr_t):
        arr = np.ascontiguousarray(arr)
        keep.append(arr)
        setattr(v, name, arr.ctypes.data_as(ptr_t))

    v.n = n
    ndr = max(int(sc.dr_start.shape[1]) for sc in scenarios)
    v.n_dr, v.lut_len = ndr, LUT_LEN

    def pad_dr(a, fill=0):
        out = np.full((Tr, ndr), fill, dtype=a.dtype)
        out[:, :a.shape[1]] = a
        return out

    for k in _lib._SCN_F64:
        put(k, np.stack([getattr(sc, k) for sc in scenarios]).astype(np.float64), _lib._pd)
    put("dr_start", np.stack([pad_dr(sc.dr_start) for sc in scenarios]).astype(np.int32), _lib._pi)
    put("dr_end", np.stack([pad_dr(sc.dr_end) for sc in scenarios]).astype(np.int32), _lib._pi)
    put("dr_cap", np.stack([pad_dr(sc.dr_cap) for sc in scenarios]).astype(np.float64), _lib._pd)
    put("dr_count", np.stack([sc.dr_count for sc in scenarios]).astype(np.int32), _lib._pi)
    s_off = np.zeros(n + 1, dtype=np.int64)
    l_off = np.zeros(n + 1, dtype=np.int64)
    for i, sc in enumerate(scenarios):
        s_off[i + 1] = s_off[i] + sc.n_sessions
        l_off[i + 1] = l_off[i] + sc.luts_c.shape[0]
    put("sess_off", s_off, _lib._pl)
    put("lut_off", l_off, _lib._pl)
    for k in SESSION_INT_FIELDS:
        put("s_" + k, np.concatenate([sc.sessions[k] for sc in scenarios]).astype(np.int32), _lib._pi)
    for k in SESSION_F64_FIELDS:
        put("s_" + k, np.concatenate([sc.sessions[k] for sc in scenarios]).astype(np.float64), _lib._pd)
    put("luts_c", np.concatenate([sc.luts_c.reshape(-1) for sc in scenarios] + [np.zeros(0)]), _lib._pd)
    put("luts_d", np.concatenate([sc.luts_d.reshape(-1) for sc in scenarios] + [np.zeros(0)]), _lib._pd)
    if all(sc.date_feat.shape == (T + 1, 3) for sc in scenarios):
        put("date_feat", np.stack([sc.date_feat for sc in scenarios]).astype(np.float64), _lib._pd)
    if topo.n_bus:
        nb = topo.n_bus
        if not all(sc.grid_active.shape == (T + 1, nb) for sc in scenarios):
            raise EngineError("grid topology: every scenario needs grid_active / grid_reactive of shape (T+1, n_bus)")
        put("grid_active", np.stack([sc.grid_active for sc in scenarios]).astype(np.float64), _lib._pd)
        put("grid_reactive", np.stack([sc.grid_reactive for sc in scenarios]).astype(np.float64), _lib._pd)
    return v, keep


class _CudaView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer zero-copy."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


class SpawnMixin:
    """Device-side scenario sampling (ev2b_set_spawn_tables / ev2b_resample_sessions / ev2b_read_sessions): shared by
    BatchedEngine and the emulator-backed test engine (same C ABI)."""

    def set_spawn_tables(self, tables):
        """Register the EV-spawner tables (scenario.SpawnTables).  Call before load_scenarios."""
        v, keep = _lib.spawn_tables_view(tables)
        self._check(self.L.ev2b_set_spawn_tables(self.h, C.byref(v)), "ev2b_set_spawn_tables")
        self._spawn_start = np.ascontiguousarray(tables.start, dtype=np.int32).reshape(-1, 3)

    def resample_sessions(self, seed: int, start=None):
        """Re-draw the EV sessions of every scenario of the bank on the device (EV_spawner, utils.py:477-557).
        start: [n_scenarios, 3] weekday / hour / minute of each scenario's sim_date (default: the tables' own, tiled)."""
        n = self.L.ev2b_n_scenarios(self.h)
        st = self._spawn_start if start is None else np.asarray(start, dtype=np.int32).reshape(-1, 3)
        if len(st) == 0:
            raise EngineError("resample_sessions needs start dates (SpawnTables.start is empty)")
        st = np.ascontiguousarray(st[np.arange(n) % len(st)], dtype=np.int32)
        self._check(self.L.ev2b_resample_sessions(self.h, int(seed) & (2 ** 64 - 1), st.ctypes.data_as(_lib._pi),
                                                  self._stream()), "ev2b_resample_sessions")

    def read_sessions(self, scn: int) -> Dict[str, np.ndarray]:
        """The sessions of scenario `scn` as the device bank holds them, in arrival order."""
        cap = self.P * 64
        i = {k: np.zeros(cap, dtype=np.int32) for k in ("port", "t_arr", "t_dep", "model")}
        d = {k: np.zeros(cap, dtype=np.float64) for k in ("cap0", "ts", "eta_c", "eta_d")}
        n = self.L.ev2b_read_sessions(self.h, int(scn), cap, *[i[k].ctypes.data_as(_lib._pi) for k in ("port", "t_arr", "t_dep", "model")],
                                      *[d[k].ctypes.data_as(_lib._pd) for k in ("cap0", "ts", "eta_c", "eta_d")])
        if n < 0:
            self._check(n, "ev2b_read_sessions")
        out = {k: v[:n].copy() for k, v in i.items()}
        out.update({k: v[:n].copy() for k, v in d.items()})
        return out

    def read_setpoints(self, scn: int) -> np.ndarray:
        """env.power_setpoints of scenario `scn` as the device bank holds them (regenerated by resample_sessions when the
        spawn tables say power_setpoint_enabled)."""
        out = np.zeros(self.topo.T, dtype=np.float64)
        self._check(self.L.ev2b_read_setpoints(self.h, int(scn), out.ctypes.data_as(_lib._pd)), "ev2b_read_setpoints")
        return out


class BatchedEngine(SpawnMixin):
    def __init__(self, topo: Topology, n_envs: int, reward=None, state=None, device: int = 0,
                 outputs: Iterable[str] = ("reward", "status", "obs"), stats: bool = False):
        import torch  # plumbing only
        if not torch.cuda.is_available():
            raise EngineError("ev2gym_b200 needs a CUDA device: there is no CPU fallback in the product path")
        self.torch = torch
        self.L = _lib.load()
        self.topo, self.E, self.device = topo, int(n_envs), int(device)
        r, s = _fn_name(reward), _fn_name(state)
        if r not in REWARD_KINDS or s not in STATE_KINDS:
            raise EngineError(f"no fused device implementation for reward={r!r} / state={s!r}")
        self.reward_name, self.state_name = r, s
        d = _lib.Dims(self.E, topo.C, topo.Tr, topo.T, topo.timescale, topo.dr_steps_ahead, REWARD_KINDS[r],
                      STATE_KINDS[s], float(topo.tr_voltage), 1 if stats else 0, 0)
        self.stats = bool(stats)
        tv, self._keep = topology_view(topo)
        h = C.c_void_p()
        rc = self.L.ev2b_create(C.byref(d), C.byref(tv), self.device, C.byref(h))
        if rc != 0:
            raise EngineError(f"ev2b_create failed ({rc}): {self.L.ev2b_last_error(None).decode()}")
        self.h = h
        self.P, self.C_, self.Tr, self.T = topo.P, topo.C, topo.Tr, topo.T
        self.D = self.L.ev2b_obs_dim(self.h)
        self.dev = torch.device("cuda", self.device)
        self.out: Dict[str, "torch.Tensor"] = {}
        self._so = _lib.StepOut()
        self.set_outputs(outputs)
        self.n_scenarios = 0

    # ------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.L.ev2b_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.ev2b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_outputs(self, names: Iterable[str]):
        torch = self.torch
        dims = {"D": max(self.D, 1), "C": self.C_, "Tr": self.Tr, "P": self.P, "N": self.topo.n_bus + 1, "T": self.T}
        self.out = {}
        self._so = _lib.StepOut()
        for n in names:
            if n == "obs" and self.D == 0:
                continue
            dt, shp = _OUT_SPECS[n]
            t = torch.zeros((self.E,) + tuple(dims[k] for k in shp), dtype=getattr(torch, dt), device=self.dev)
            self.out[n] = t
            setattr(self._so, n, t.data_ptr())

    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    # ------------------------------------------------------------------------------------------
    def load_scenarios(self, scenarios: Sequence[Scenario]):
        """Upload a scenario bank (replaces the previous one; every env must be reset afterwards)."""
        v, _keep = scenarios_view(self.topo, scenarios)
        self._check(self.L.ev2b_load_scenarios(self.h, C.byref(v)), "ev2b_load_scenarios")
        self.n_scenarios = len(scenarios)

    def reset(self, env_lo: int = 0, env_hi: Optional[int] = None, scn_ids: Optional[Sequence[int]] = None):
        """Reset envs [env_lo, env_hi); returns the observation tensor [E,D] (rows of that range refreshed)."""
        env_hi = self.E if env_hi is None else env_hi
        ids = None
        if scn_ids is not None:
            ids_np = np.ascontiguousarray(scn_ids, dtype=np.int32)
            assert ids_np.shape == (env_hi - env_lo,)
            ids = ids_np.ctypes.data_as(_lib._pi)
        obs = self.out.get("obs")
        self._check(self.L.ev2b_reset(self.h, env_lo, env_hi, ids, obs.data_ptr() if obs is not None else None,
                                      self._stream()), "ev2b_reset")
        return obs

    def reset_done(self):
        obs = self.out.get("obs")
        self._check(self.L.ev2b_reset_done(self.h, obs.data_ptr() if obs is not None else None, self._stream()),
                    "ev2b_reset_done")
        return obs

    def step(self, actions) -> Dict[str, "torch.Tensor"]:
        """One fused kernel launch advancing all E envs.  actions: cuda tensor [E,P], float32 or float64."""
        torch = self.torch
        if actions.device != self.dev or tuple(actions.shape) != (self.E, self.P) or not actions.is_contiguous():
            raise EngineError(f"actions must be a contiguous cuda tensor of shape {(self.E, self.P)} on {self.dev}")
        if actions.dtype == torch.float32:
            dt = 0
        elif actions.dtype == torch.float64:
            dt = 1
        else:
            raise EngineError("actions must be float32 or float64")
        self._check(self.L.ev2b_step(self.h, actions.data_ptr(), dt, C.byref(self._so), self._stream()), "ev2b_step")
        return self.out

    AGENTS = {"external": 0, "afap": 1, "zero": 2, "uniform": 3, "roundrobin": 4, "calap": 5}

    def agent_actions(self, agent: str, out=None):
        """`agent.get_action(env)` of a stock heuristic (ev2gym/baselines/heuristics.py) for every env's current state:
        float64 cuda tensor [E,P] to pass to `step`.  "afap", "zero", "roundrobin" (stateful: one call == one
        get_action; the per-env queue empties when the env is reset), "calap"."""
        torch = self.torch
        if out is None:
            out = torch.empty((self.E, self.P), dtype=torch.float64, device=self.dev)
        if out.dtype != torch.float64 or tuple(out.shape) != (self.E, self.P) or not out.is_contiguous() or not out.is_cuda:
            raise EngineError(f"out must be a contiguous float64 cuda tensor of shape {(self.E, self.P)}")
        self._check(self.L.ev2b_agent_actions(self.h, self.AGENTS[agent], out.data_ptr(), self._stream()),
                    "ev2b_agent_actions")
        return out

    def step_k(self, k: int, agent: str = "afap", actions_k=None, seed: int = 0, auto_reset: bool = False):
        """k steps without returning to the host, driven by an on-device agent (or actions_k [k,E,P])."""
        kind = self.AGENTS[agent]
        dt, ptr = 0, None
        if kind == 0:
            if tuple(actions_k.shape) != (k, self.E, self.P) or not actions_k.is_contiguous():
                raise EngineError(f"actions_k must be a contiguous cuda tensor of shape {(k, self.E, self.P)}")
            dt = 1 if actions_k.dtype == self.torch.float64 else 0
            ptr = actions_k.data_ptr()
        low = -1.0 if self.topo.v2g_enabled else 0.0
        self._check(self.L.ev2b_step_k(self.h, int(k), kind, ptr, dt, int(seed) & (2 ** 64 - 1), low, int(auto_reset),
                                       C.byref(self._so), self._stream()), "ev2b_step_k")
        return self.out

    @staticmethod
    def uniform_agent_actions(seed: int, n_envs: int, n_ports: int, t: int, low: float) -> np.ndarray:
        """Host mirror of the UNIFORM device agent (include/ev2b.h): exactly the float64 actions it draws."""
        def mix32(x):
            x = x.astype(np.uint64) & 0xFFFFFFFF
            x ^= x >> 16; x = (x * 0x7feb352d) & 0xFFFFFFFF; x ^= x >> 15
            x = (x * 0x846ca68b) & 0xFFFFFFFF; x ^= x >> 16
            return x
        ip = np.arange(n_envs * n_ports, dtype=np.uint64)
        lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
        h = mix32((mix32(ip ^ lo) + np.uint64((t * 0x9E3779B9) & 0xFFFFFFFF) + hi) & 0xFFFFFFFF)
        u = (h >> 8).astype(np.float64) * (1.0 / 16777216.0)
        return (low + (1.0 - low) * u).reshape(n_envs, n_ports)

    def step_host(self, actions: np.ndarray, reward: np.ndarray, status: np.ndarray, obs: Optional[np.ndarray] = None):
        """End-to-end step with HOST buffers (H2D actions, kernel, D2H reward/status[/obs], sync)."""
        dt = {np.dtype("float32"): 0, np.dtype("float64"): 1}[actions.dtype]
        assert actions.shape == (self.E, self.P) and actions.flags.c_contiguous
        self._check(self.L.ev2b_step_host(self.h, actions.ctypes.data, dt, reward.ctypes.data, status.ctypes.data,
                                          obs.ctypes.data if obs is not None else None, self._stream()),
                    "ev2b_step_host")

    # ------------------------------------------------------------------------------------------
    def state_tensors(self) -> Dict[str, "torch.Tensor"]:
        """Zero-copy torch views of the SoA state (valid until the engine is closed)."""
        torch = self.torch
        sv = _lib.StateView()
        self._check(self.L.ev2b_state_view_get(self.h, C.byref(sv)), "ev2b_state_view_get")
        E, P = self.E, self.P

        def view(ptr, shape, typestr):
            return torch.as_tensor(_CudaView(ptr, shape, typestr, self), device=self.dev)
        return {
            "port_cap": view(sv.port_cap, (E, P), "<f8"), "port_exch": view(sv.port_exch, (E, P), "<f8"),
            "port_hot": view(sv.port_hot, (E, P, 4), "<i4"), "env_step": view(sv.env_step, (E,), "<i4"),
            "env_scn": view(sv.env_scn, (E,), "<i4"), "env_potential": view(sv.env_potential, (E,), "<f8"),
            "env_usage": view(sv.env_usage, (E,), "<f8"), "env_kpi": view(sv.env_kpi, (E, sv.n_kpi), "<f8"),
        }

    def episode_stats(self) -> Dict[str, np.ndarray]:
        """get_statistics(env) of every env (ev2gym/utilities/utils.py:12-123); needs stats=True."""
        torch = self.torch
        out = torch.zeros((self.E, len(STAT_NAMES)), dtype=torch.float64, device=self.dev)
        self._check(self.L.ev2b_episode_stats(self.h, out.data_ptr(), self._stream()), "ev2b_episode_stats")
        o = out.cpu().numpy()
        return {n: o[:, i].copy() for i, n in enumerate(STAT_NAMES)}

    def histories(self) -> Dict[str, "torch.Tensor"]:
        """The history outputs that were asked for (set_outputs(..., "hist_cs_power", ...)) in the reference's shapes:
        cs_power / cs_current [E,C,T], tr_overload [E,Tr,T], current_power_usage [E,T]  (utils.py:794-861)."""
        names = {"hist_cs_power": "cs_power", "hist_cs_current": "cs_current", "hist_tr_overload": "tr_overload",
                 "hist_usage": "current_power_usage"}
        return {v: (self.out[k].transpose(1, 2) if self.out[k].dim() == 3 else self.out[k])
                for k, v in names.items() if k in self.out}

    def kpis(self) -> Dict[str, np.ndarray]:
        k = self.state_tensors()["env_kpi"].cpu().numpy()
        return {n: k[:, i].copy() for i, n in enumerate(KPI_NAMES)}

    @property
    def launch_count(self) -> int:
        return int(self.L.ev2b_launch_count(self.h))

    def kernel_launches(self):
        """Step launches by kernel: (step_kernel, evl_step_kernel, evl_rebuild_kernel)  -- include/ev2b.h."""
        return tuple(int(self.L.ev2b_kernel_launches(self.h, k)) for k in range(3))

    @staticmethod
    def decode_hot(hot: np.ndarray) -> Dict[str, np.ndarray]:
        """Unpack the [.,4] int32 hot words into t_arr / t_dep / next_arr / spec fields."""
        w = hot.astype(np.int64) & 0xFFFFFFFF
        i16 = lambda x: ((x & 0xFFFF) ^ 0x8000) - 0x8000
        return {"t_arr": i16(w[..., 0]), "t_dep": i16(w[..., 0] >> 16), "next_arr": i16(w[..., 1]),
                "cursor": (w[..., 1] >> 16) & 0xFF, "spec": w[..., 2] & 0xFFFF, "ts_milli": w[..., 2] >> 16}
