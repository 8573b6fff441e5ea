"""ctypes front-end of the C ORACLE (oracle/ev2o.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; nothing under ev2gym_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libev2oracle.so")

REWARD_KINDS = {None: 0, "none": 0, "SquaredTrackingErrorReward": 1,
                "ProfitMax_TrPenalty_UserIncentives": 2, "profit_maximization": 3, "V2G_grid_full_reward": 4,
                "V2G_grid_simple_reward": 5, "SqTrError_TrPenalty_UserIncentives": 6, "SimpleReward": 7,
                "MinimizeTrackerSurplusWithChargeRewards": 8, "V2G_profitmax": 9, "V2G_costs_simple": 10,
                "V2G_profitmaxV2": 11, "Grid_V2G_profitmaxV2": 12, "pst_V2G_profitmaxV2": 13,
                "SquaredTrackingErrorRewardWithPenalty": 14}
STATE_KINDS = {None: 0, "none": 0, "PublicPST": 1, "V2G_profit_max": 2, "V2G_profit_max_loads": 3, "V2G_grid_state": 4}

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


class _Topo(C.Structure):
    _fields_ = [("C", C.c_int), ("P", C.c_int), ("Tr", C.c_int), ("T", C.c_int), ("timescale", C.c_int),
                ("dr_steps_ahead", C.c_int),
                ("cs_n_ports", _pi), ("cs_port_off", _pi), ("cs_tr", _pi), ("cs_phases", _pi),
                ("cs_imax", _pd), ("cs_imin", _pd), ("cs_imax_dis", _pd), ("cs_imin_dis", _pd),
                ("cs_voltage", _pd), ("tr_voltage", C.c_double),
                ("n_bus", C.c_int), ("grid_K", _pd), ("grid_L", _pd), ("s_base", C.c_double)]


class _Scn(C.Structure):
    _fields_ = [("charge_price", _pd), ("discharge_price", _pd), ("setpoint", _pd),
                ("tr_infl", _pd), ("tr_solar", _pd), ("tr_max_power", _pd), ("tr_min_power", _pd),
                ("tr_load_fc", _pd), ("tr_pv_fc", _pd),
                ("n_dr", C.c_int), ("dr_start", _pi), ("dr_end", _pi), ("dr_count", _pi), ("dr_cap", _pd),
                ("n_sessions", C.c_int),
                ("s_loc", _pi), ("s_t_arr", _pi), ("s_t_dep", _pi), ("s_ev_phases", _pi), ("s_lut", _pi),
                ("s_cap0", _pd), ("s_B", _pd), ("s_pmax_ac", _pd), ("s_pmin_ac", _pd), ("s_pmax_dis", _pd),
                ("s_pmin_dis", _pd), ("s_bmin", _pd), ("s_bmin_em", _pd), ("s_desired", _pd), ("s_ts", _pd),
                ("s_mult", _pd), ("s_eta_c", _pd), ("s_eta_d", _pd),
                ("n_luts", C.c_int), ("lut_len", C.c_int), ("luts_c", _pd), ("luts_d", _pd),
                ("grid_active", _pd), ("grid_reactive", _pd), ("date_feat", _pd)]


_STATE_ARRAYS = [("port_session", "i", "P"), ("port_cap", "d", "P"), ("port_energy_exch", "d", "P"),
                 ("port_abs_energy", "d", "P"), ("port_prev_power", "d", "P"), ("port_required", "d", "P"),
                 ("port_cur_energy", "d", "P"), ("port_cur_amps", "d", "P"), ("port_cycles", "i", "P"),
                 ("port_em_metric", "i", "P"),
                 ("cs_total_charged", "d", "C"), ("cs_total_discharged", "d", "C"), ("cs_total_profits", "d", "C"),
                 ("cs_total_sat", "d", "C"), ("cs_total_served", "i", "C"),
                 ("usage", "d", "T"), ("potential", "d", "T"), ("tr_overload_hist", "d", "TrT"),
                 ("cs_power_hist", "d", "CT"), ("cs_current_hist", "d", "CT"), ("node_voltage", "d", "NT"),
                 ("load_fc_live", "d", "TrT"), ("pv_fc_live", "d", "TrT"),
                 ("ev_spawned", "i", "S"), ("ev_final_cap", "d", "S"), ("ev_afap", "d", "S"), ("ev_soc_sum", "d", "S"),
                 ("ev_n_hist", "i", "S"), ("ev_abs_energy", "d", "S"), ("ev_em_metric", "i", "S"), ("ev_n_act", "i", "S"),
                 ("ev_act_soc", "d", "ST")]


class _State(C.Structure):
    _fields_ = [("current_step", C.c_int), ("total_evs_spawned", C.c_int), ("current_evs_parked", C.c_int),
                ("done", C.c_int), ("total_reward", C.c_double)] + \
               [(n, _pi if k == "i" else _pd) for n, k, _ in _STATE_ARRAYS]


_OUT_ARRAYS = [("cs_power", "C"), ("cs_current", "C"), ("tr_power", "Tr"), ("tr_amps", "Tr"),
               ("tr_overload", "Tr"), ("dep_sat", "P"), ("action_mask", "P"), ("obs", "D"), ("actions_eff", "P"),
               ("node_vm", "N")]


class _Out(C.Structure):
    _fields_ = [("reward", C.c_double), ("total_costs", C.c_double), ("done", C.c_int),
                ("invalid_actions", C.c_int), ("n_departed", C.c_int), ("n_arrived", C.c_int), ("error", C.c_int)] + \
               [(n, _pd) for n, _ in _OUT_ARRAYS] + [("pf_iterations", C.c_int)]


def build(force: bool = False) -> str:
    """Compile oracle/libev2oracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("ev2o.c", "ev2o_batch.c", "ev2o.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or \
            any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ev2o_obs_dim.restype = C.c_int
        L.ev2o_obs_dim.argtypes = [C.POINTER(_Topo), C.c_int]
        L.ev2o_reset.restype = None
        L.ev2o_reset.argtypes = [C.POINTER(_Topo), C.POINTER(_Scn), C.POINTER(_State), C.c_int, _pd]
        L.ev2o_step.restype = C.c_int
        L.ev2o_step.argtypes = [C.POINTER(_Topo), C.POINTER(_Scn), C.POINTER(_State), _pd, C.c_int, C.c_int,
                                C.POINTER(_Out)]
        L.ev2o_ev_step.restype = C.c_double
        L.ev2o_ev_step.argtypes = [_pd, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _pd, _pd]
        L.ev2o_statistics.restype = None
        L.ev2o_statistics.argtypes = [C.POINTER(_Topo), C.POINTER(_Scn), C.POINTER(_State), _pd]
        L.ev2o_max_threads.restype = C.c_int
        L.ev2o_step_batch.restype = C.c_int
        L.ev2o_step_batch.argtypes = [C.POINTER(_Topo), C.POINTER(C.POINTER(_Scn)), C.POINTER(_State), C.c_int,
                                      _pd, C.c_int, C.c_int, C.POINTER(_Out), _pd, _pi, C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    if a.dtype == np.float64:
        return a.ctypes.data_as(_pd)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_pi)
    raise TypeError(a.dtype)


class _TopoC:
    def __init__(self, topo):
        self.topo = topo
        self.port_off = np.ascontiguousarray(topo.cs_port_off, dtype=np.int32)
        t = _Topo()
        t.C, t.P, t.Tr, t.T, t.timescale, t.dr_steps_ahead = topo.C, topo.P, topo.Tr, topo.T, topo.timescale, \
            topo.dr_steps_ahead
        t.cs_n_ports, t.cs_port_off, t.cs_tr, t.cs_phases = _p(topo.cs_n_ports), _p(self.port_off), \
            _p(topo.cs_tr), _p(topo.cs_phases)
        t.cs_imax, t.cs_imin, t.cs_imax_dis, t.cs_imin_dis, t.cs_voltage = _p(topo.cs_imax), _p(topo.cs_imin), \
            _p(topo.cs_imax_dis), _p(topo.cs_imin_dis), _p(topo.cs_voltage)
        t.tr_voltage = float(topo.tr_voltage)
        t.n_bus = topo.n_bus
        if topo.n_bus:
            self._K = np.ascontiguousarray(topo.grid_K).view(np.float64).reshape(-1)
            self._L = np.ascontiguousarray(topo.grid_L).view(np.float64).reshape(-1)
            t.grid_K, t.grid_L, t.s_base = _p(self._K), _p(self._L), float(topo.grid_s_base)
        self.c = t


class _ScnC:
    def __init__(self, sc):
        sc.normalise()
        self.sc = sc  # keeps the numpy buffers alive
        s = _Scn()
        for k in ("charge_price", "discharge_price", "setpoint", "tr_infl", "tr_solar", "tr_max_power",
                  "tr_min_power", "tr_load_fc", "tr_pv_fc", "dr_start", "dr_end", "dr_count", "dr_cap"):
            setattr(s, k, _p(getattr(sc, k)))
        s.n_dr = int(sc.dr_start.shape[1]) if sc.dr_start.ndim == 2 else 0
        s.n_sessions = sc.n_sessions
        for k, v in sc.sessions.items():
            if hasattr(s, "s_" + k):
                setattr(s, "s_" + k, _p(v))
        s.n_luts, s.lut_len = sc.luts_c.shape[0], sc.luts_c.shape[1]
        s.luts_c, s.luts_d = _p(sc.luts_c), _p(sc.luts_d)
        self._ga = np.ascontiguousarray(sc.grid_active, dtype=np.float64).reshape(-1)
        self._gr = np.ascontiguousarray(sc.grid_reactive, dtype=np.float64).reshape(-1)
        self._df = np.ascontiguousarray(sc.date_feat, dtype=np.float64).reshape(-1)
        if self._ga.size:
            s.grid_active, s.grid_reactive = _p(self._ga), _p(self._gr)
        if self._df.size:
            s.date_feat = _p(self._df)
        self.c = s


STAT_NAMES = ("total_ev_served", "total_profits", "total_energy_charged", "total_energy_discharged",
              "average_user_satisfaction", "power_tracker_violation", "tracking_error", "energy_tracking_error",
              "energy_user_satisfaction", "std_energy_user_satisfaction", "min_energy_user_satisfaction",
              "total_steps_min_emergency_battery_capacity_violation", "total_transformer_overload",
              "battery_degradation", "battery_degradation_calendar", "battery_degradation_cycling", "total_reward")


def _sizes(topo, D, S=1):
    S = max(int(S), 1)
    return {"N": topo.n_bus + 1, "NT": (topo.n_bus + 1) * topo.T, "S": S, "ST": S * topo.T, "P": topo.P, "C": topo.C, "T": topo.T, "Tr": max(topo.Tr, 1), "TrT": max(topo.Tr, 1) * topo.T,
            "CT": topo.C * topo.T, "D": max(D, 1)}


class OracleEnv:
    """One reference-equivalent env: reset() / step(actions) with every per-step quantity exposed."""

    def __init__(self, topo, scenario, reward: Optional[str] = None, state: Optional[str] = None):
        self.L = lib()
        self.topo, self.scenario = topo, scenario
        self.reward_kind, self.state_kind = REWARD_KINDS[reward], STATE_KINDS[state]
        self._t, self._s = _TopoC(topo), _ScnC(scenario)
        self.obs_dim = self.L.ev2o_obs_dim(C.byref(self._t.c), self.state_kind)
        sz = _sizes(topo, self.obs_dim, scenario.n_sessions)
        self.state = _State()
        self.arr = {}
        for n, k, dim in _STATE_ARRAYS:
            a = np.zeros(sz[dim], dtype=np.int32 if k == "i" else np.float64)
            self.arr[n] = a
            setattr(self.state, n, _p(a))
        self.out = _Out()
        self.o = {}
        for n, dim in _OUT_ARRAYS:
            a = np.zeros(sz[dim], dtype=np.float64)
            self.o[n] = a
            setattr(self.out, n, _p(a))

    def reset(self) -> np.ndarray:
        self.L.ev2o_reset(C.byref(self._t.c), C.byref(self._s.c), C.byref(self.state), self.state_kind,
                          _p(self.o["obs"]))
        return self.o["obs"][:self.obs_dim].copy()

    def step(self, actions: Sequence[float]) -> dict:
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.shape == (self.topo.P,)
        rc = self.L.ev2o_step(C.byref(self._t.c), C.byref(self._s.c), C.byref(self.state), _p(a),
                              self.reward_kind, self.state_kind, C.byref(self.out))
        if rc == 2:
            raise AssertionError("Episode is done, please reset the environment")
        r = {n: self.o[n].copy() for n, _ in _OUT_ARRAYS}
        r["obs"] = r["obs"][:self.obs_dim]
        r.update(reward=self.out.reward, total_costs=self.out.total_costs, done=bool(self.out.done),
                 invalid_actions=self.out.invalid_actions, n_departed=self.out.n_departed,
                 n_arrived=self.out.n_arrived, error=self.out.error,
                 port_cap=self.arr["port_cap"].copy(), port_session=self.arr["port_session"].copy(),
                 port_energy_exch=self.arr["port_energy_exch"].copy(),
                 current_step=self.state.current_step, potential=self.arr["potential"].copy(),
                 usage=self.arr["usage"].copy())
        return r

    def statistics(self) -> dict:
        out = np.zeros(len(STAT_NAMES))
        self.L.ev2o_statistics(C.byref(self._t.c), C.byref(self._s.c), C.byref(self.state), _p(out))
        return dict(zip(STAT_NAMES, out.tolist()))

    @property
    def current_step(self) -> int:
        return self.state.current_step

    @property
    def total_reward(self) -> float:
        return self.state.total_reward


class OracleBatch:
    """E independent oracle envs stepped together on host threads (CPU baseline / `--impl reference`)."""

    def __init__(self, topo, scenarios: List, reward: Optional[str] = None, state: Optional[str] = None,
                 threads: int = 0):
        self.L = lib()
        self.topo = topo
        self.E = len(scenarios)
        self.threads = threads if threads > 0 else self.L.ev2o_max_threads()
        self.reward_kind, self.state_kind = REWARD_KINDS[reward], STATE_KINDS[state]
        self._t = _TopoC(topo)
        self._scenarios = list(scenarios)
        uniq = {}
        self._scn = []
        for sc in scenarios:           # scenarios may repeat (tiling): share the C view
            if id(sc) not in uniq:
                uniq[id(sc)] = _ScnC(sc)
            self._scn.append(uniq[id(sc)])
        self._scn_ptrs = (C.POINTER(_Scn) * self.E)(*[C.pointer(s.c) for s in self._scn])
        self.obs_dim = self.L.ev2o_obs_dim(C.byref(self._t.c), self.state_kind)
        sz = _sizes(topo, self.obs_dim, max(sc.n_sessions for sc in scenarios))
        self.states = (_State * self.E)()
        self.outs = (_Out * self.E)()
        self.arr = {}
        for n, k, dim in _STATE_ARRAYS:
            a = np.zeros((self.E, sz[dim]), dtype=np.int32 if k == "i" else np.float64)
            self.arr[n] = a
            for e in range(self.E):
                setattr(self.states[e], n, _p(a[e]))
        self.o = {}
        for n, dim in _OUT_ARRAYS:
            a = np.zeros((self.E, sz[dim]), dtype=np.float64)
            self.o[n] = a
            for e in range(self.E):
                setattr(self.outs[e], n, _p(a[e]))
        self.reward = np.zeros(self.E)
        self.done = np.zeros(self.E, dtype=np.int32)

    def reset(self):
        for e in range(self.E):
            self.L.ev2o_reset(C.byref(self._t.c), self._scn_ptrs[e], C.byref(self.states[e]), self.state_kind,
                              _p(self.o["obs"][e]))
        return self.o["obs"][:, :self.obs_dim]

    def statistics(self) -> dict:
        out = np.zeros((self.E, len(STAT_NAMES)))
        for e in range(self.E):
            self.L.ev2o_statistics(C.byref(self._t.c), self._scn_ptrs[e], C.byref(self.states[e]), _p(out[e]))
        return {n: out[:, i].copy() for i, n in enumerate(STAT_NAMES)}

    def env_view(self, e: int):
        """What oracle/agents.py reads of one env (same attribute names as OracleEnv), live views of env `e`."""
        batch = self

        class _View:
            topo = batch.topo
            scenario = batch._scenarios[e]
            arr = {k: v[e] for k, v in batch.arr.items()}

            @property
            def current_step(self):
                return batch.states[e].current_step
        return _View()

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.shape == (self.E, self.topo.P)
        rc = self.L.ev2o_step_batch(C.byref(self._t.c), self._scn_ptrs, self.states, self.E, _p(a.reshape(-1)),
                                    self.reward_kind, self.state_kind, self.outs, _p(self.reward), _p(self.done),
                                    self.threads)
        return self.reward, self.done, rc
