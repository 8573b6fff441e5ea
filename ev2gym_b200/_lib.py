"""Build and load libev2b.so (the C-ABI CUDA library, include/ev2b.h) through ctypes.

There is NO CPU fallback: if the library is missing or no CUDA device is present the engine
raises.  The library is built in-tree (ev2gym_b200/csrc/libev2b.so) so it travels with the repo.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("EV2B_LIB") or os.path.join(CSRC, "libev2b.so")   # EV2B_LIB: A/B-test another build
SOURCES = [os.path.join(CSRC, f) for f in ("ev2b.cu", "ev2b_device.cuh", "ev2b_evlist.cuh", "ev2b_spawn.cuh", "ev2b_math.h")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "ev2b.h")]

# No -split-compile: nvcc's parallel optimisation splits the translation unit differently from run to run (same sources,
# same flags: 14.6 or 17.8 MB of cubin, 48 or 120 bytes of spill stack in evl_step_kernel), and the step kernel's time
# moved by +-5 % with it on the B200 (profiles/r2_ab_build_modes.jsonl).  One thread: ~3 min, the same binary every time.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off"]

_pd, _pi, _pf = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_float)
_pl, _pu, _pb = C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)


class Dims(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("n_chargers", C.c_int32), ("n_transformers", C.c_int32),
                ("sim_length", C.c_int32), ("timescale", C.c_int32), ("dr_steps_ahead", C.c_int32),
                ("reward_kind", C.c_int32), ("state_kind", C.c_int32), ("tr_voltage", C.c_double),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class TopologyView(C.Structure):
    _fields_ = [("cs_n_ports", _pi), ("cs_tr", _pi), ("cs_phases", _pi), ("cs_imax", _pd), ("cs_imin", _pd),
                ("cs_imax_dis", _pd), ("cs_imin_dis", _pd), ("cs_voltage", _pd),
                ("n_bus", C.c_int32), ("grid_K", _pd), ("grid_L", _pd), ("grid_s_base", C.c_double)]


_SCN_F64 = ("charge_price", "discharge_price", "setpoint", "tr_infl", "tr_solar", "tr_max_power", "tr_min_power",
            "tr_load_fc", "tr_pv_fc")
_SESS_I = ("s_loc", "s_t_arr", "s_t_dep", "s_ev_phases", "s_lut")
_SESS_D = ("s_cap0", "s_B", "s_pmax_ac", "s_pmin_ac", "s_pmax_dis", "s_pmin_dis", "s_bmin", "s_bmin_em",
           "s_desired", "s_ts", "s_mult", "s_eta_c", "s_eta_d")


class ScenariosView(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_dr", C.c_int32), ("lut_len", C.c_int32)] + \
               [(k, _pd) for k in _SCN_F64] + \
               [("dr_start", _pi), ("dr_end", _pi), ("dr_cap", _pd), ("dr_count", _pi), ("sess_off", _pl)] + \
               [(k, _pi) for k in _SESS_I] + [(k, _pd) for k in _SESS_D] + \
               [("lut_off", _pl), ("luts_c", _pd), ("luts_d", _pd), ("grid_active", _pd), ("grid_reactive", _pd),
                ("date_feat", _pd)]


class StepOut(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("status", C.c_void_p), ("obs", C.c_void_p), ("cs_power", C.c_void_p),
                ("cs_current", C.c_void_p), ("tr_power", C.c_void_p), ("tr_overload", C.c_void_p),
                ("total_costs", C.c_void_p), ("action_mask", C.c_void_p), ("dep_sat", C.c_void_p),
                ("dep_cap", C.c_void_p), ("port_energy", C.c_void_p), ("node_voltage", C.c_void_p),
                ("hist_cs_power", C.c_void_p), ("hist_cs_current", C.c_void_p), ("hist_tr_overload", C.c_void_p),
                ("hist_usage", C.c_void_p)]


class StateView(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("n_ports", C.c_int32), ("n_chargers", C.c_int32),
                ("n_transformers", C.c_int32), ("obs_dim", C.c_int32), ("n_kpi", C.c_int32),
                ("port_cap", C.c_void_p), ("port_exch", C.c_void_p), ("port_hot", C.c_void_p),
                ("env_step", C.c_void_p), ("env_scn", C.c_void_p), ("env_potential", C.c_void_p),
                ("env_usage", C.c_void_p), ("env_kpi", C.c_void_p)]


ABI_VERSION = 2          # EV2B_ABI_VERSION of include/ev2b.h this binding was written against

EXPORTS = ("ev2b_abi_version", "ev2b_last_error", "ev2b_create", "ev2b_destroy", "ev2b_obs_dim", "ev2b_n_ports",
           "ev2b_load_scenarios", "ev2b_n_scenarios", "ev2b_reset", "ev2b_step", "ev2b_step_host",
           "ev2b_reset_done", "ev2b_state_view_get", "ev2b_launch_count", "ev2b_episode_stats", "ev2b_step_k",
           "ev2b_agent_actions", "ev2b_kernel_launches", "ev2b_set_spawn_tables", "ev2b_resample_sessions",
           "ev2b_read_sessions", "ev2b_read_setpoints")


class SpawnTablesView(C.Structure):
    _fields_ = [("workplace", C.c_int32), ("heterogeneous", C.c_int32), ("empty_ports_at_end", C.c_int32),
                ("min_stay_steps", C.c_int32), ("n_models", C.c_int32), ("n_luts", C.c_int32),
                ("power_setpoint_enabled", C.c_int32), ("reserved", C.c_int32),
                ("spawn_multiplier", C.c_double), ("desired_frac", C.c_double), ("min_battery_capacity", C.c_double),
                ("min_emergency_battery_capacity", C.c_double), ("ts_multiplier", C.c_double),
                ("homog_ts", C.c_double), ("homog_eta_c", C.c_double), ("homog_eta_d", C.c_double),
                ("power_setpoint_flexibility", C.c_double), ("arrival_week", _pd), ("arrival_weekend", _pd), ("req_energy_mean", _pd), ("stay_mean", _pd),
                ("model_prob", _pd), ("model_B", _pd), ("model_pmax_ac", _pd), ("model_pmax_dis", _pd),
                ("model_pmin_ac", _pd), ("model_pmin_dis", _pd), ("model_phases", _pi), ("model_lut", _pi), ("luts", _pd)]


def spawn_tables_view(t):
    """`ev2b_spawn_tables` over a scenario.SpawnTables; returns (view, arrays to keep alive)."""
    import numpy as np
    v, keep = SpawnTablesView(), []
    for k in ("workplace", "heterogeneous", "empty_ports_at_end", "min_stay_steps", "power_setpoint_enabled"):
        setattr(v, k, int(getattr(t, k)))
    for k in ("spawn_multiplier", "desired_frac", "min_battery_capacity", "min_emergency_battery_capacity",
              "ts_multiplier", "homog_ts", "homog_eta_c", "homog_eta_d", "power_setpoint_flexibility"):
        setattr(v, k, float(getattr(t, k)))
    v.n_models, v.n_luts = int(len(t.model_prob)), int(np.asarray(t.luts).reshape(-1, 101).shape[0])
    for k, ptr, dt in [("arrival_week", _pd, np.float64), ("arrival_weekend", _pd, np.float64),
                       ("req_energy_mean", _pd, np.float64), ("stay_mean", _pd, np.float64), ("model_prob", _pd, np.float64),
                       ("model_B", _pd, np.float64), ("model_pmax_ac", _pd, np.float64), ("model_pmax_dis", _pd, np.float64),
                       ("model_pmin_ac", _pd, np.float64), ("model_pmin_dis", _pd, np.float64),
                       ("model_phases", _pi, np.int32), ("model_lut", _pi, np.int32), ("luts", _pd, np.float64)]:
        a = np.ascontiguousarray(np.asarray(getattr(t, k)).reshape(-1), dtype=dt)
        if a.size == 0:
            a = np.zeros(1, dtype=dt)
        keep.append(a)
        setattr(v, k, a.ctypes.data_as(ptr))
    return v, keep


def declare_spawn(L):
    L.ev2b_set_spawn_tables.restype = C.c_int
    L.ev2b_set_spawn_tables.argtypes = [C.c_void_p, C.POINTER(SpawnTablesView)]
    L.ev2b_resample_sessions.restype = C.c_int
    L.ev2b_resample_sessions.argtypes = [C.c_void_p, C.c_uint64, _pi, C.c_void_p]
    L.ev2b_read_sessions.restype = C.c_int
    L.ev2b_read_setpoints.restype = C.c_int
    L.ev2b_read_setpoints.argtypes = [C.c_void_p, C.c_int, _pd]
    L.ev2b_read_sessions.argtypes = [C.c_void_p, C.c_int, C.c_int, _pi, _pi, _pi, _pi, _pd, _pd, _pd, _pd]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    if os.environ.get("EV2B_LIB"):       # an explicitly chosen build is used as it is, never rebuilt from the sources
        return False
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compile libev2b.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and ev2gym_b200/csrc/libev2b.so is missing or stale")
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:       # several ranks may find the library stale at once: one builds, the
        fcntl.flock(lock, fcntl.LOCK_EX)             # others wait and then find it fresh; the .so appears atomically
        if force or needs_build():
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC, "ev2b.cu")]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.check_call(cmd)
            os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None


def load():
    """Load libev2b.so; build it first if sources are newer and nvcc exists.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if os.path.exists(nvcc):
            build()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError("libev2b.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    L = C.CDLL(LIB_PATH)
    L.ev2b_abi_version.restype = C.c_int
    L.ev2b_last_error.restype = C.c_char_p
    L.ev2b_last_error.argtypes = [C.c_void_p]
    L.ev2b_create.restype = C.c_int
    L.ev2b_create.argtypes = [C.POINTER(Dims), C.POINTER(TopologyView), C.c_int, C.POINTER(C.c_void_p)]
    L.ev2b_destroy.restype = None
    L.ev2b_destroy.argtypes = [C.c_void_p]
    for f in ("ev2b_obs_dim", "ev2b_n_ports", "ev2b_n_scenarios"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [C.c_void_p]
    L.ev2b_load_scenarios.restype = C.c_int
    L.ev2b_load_scenarios.argtypes = [C.c_void_p, C.POINTER(ScenariosView)]
    L.ev2b_reset.restype = C.c_int
    L.ev2b_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, _pi, C.c_void_p, C.c_void_p]
    L.ev2b_step.restype = C.c_int
    L.ev2b_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(StepOut), C.c_void_p]
    L.ev2b_step_host.restype = C.c_int
    L.ev2b_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ev2b_reset_done.restype = C.c_int
    L.ev2b_reset_done.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ev2b_state_view_get.restype = C.c_int
    L.ev2b_state_view_get.argtypes = [C.c_void_p, C.POINTER(StateView)]
    L.ev2b_agent_actions.restype = C.c_int
    L.ev2b_agent_actions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ev2b_step_k.restype = C.c_int
    L.ev2b_step_k.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_int,
                              C.POINTER(StepOut), C.c_void_p]
    L.ev2b_episode_stats.restype = C.c_int
    L.ev2b_episode_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ev2b_launch_count.restype = C.c_int64
    L.ev2b_launch_count.argtypes = [C.c_void_p]
    L.ev2b_kernel_launches.restype = C.c_int64
    L.ev2b_kernel_launches.argtypes = [C.c_void_p, C.c_int]
    declare_spawn(L)
    if L.ev2b_abi_version() != ABI_VERSION:
        raise RuntimeError("libev2b.so ABI version mismatch")
    _lib = L
    return L
