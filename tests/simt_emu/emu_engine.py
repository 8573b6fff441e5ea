"""TEST INFRASTRUCTURE ONLY: drives libev2b_emu.so -- ev2gym_b200/csrc/ev2b.cu compiled by g++ against the SIMT
emulator (simt_emu.h) -- through the same C ABI as libev2b.so, with numpy arrays standing in for device memory.
Nothing under ev2gym_b200/ imports this; the product library has no CPU path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_BUILD = os.path.join(_HERE, "_build")
_EXTRA = os.environ.get("EV2B_EMU_CXXFLAGS", "").split()          # e.g. -DEV2B_EVL_PIPELINE=1: emulate a build variant
LIB_PATH = os.path.join(_BUILD, "libev2b_emu%s.so" % ("_" + "".join(c for c in "".join(_EXTRA) if c.isalnum()) if _EXTRA else ""))
_CSRC = os.path.join(_ROOT, "ev2gym_b200", "csrc")
_SOURCES = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(_HERE, f) for f in ("simt_emu.h", "simt_emu.cc")] + [os.path.join(_ROOT, "include", "ev2b.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in _SOURCES):
        return LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:          # pytest-xdist workers: one builds, the others wait and re-check
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in _SOURCES):
            return LIB_PATH
        tmp = LIB_PATH + ".tmp%d" % os.getpid()
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-DEV2B_SIMT_EMU", "-I", _HERE] + _EXTRA + [
               "-x", "c++", os.path.join(_CSRC, "ev2b.cu"), os.path.join(_HERE, "simt_emu.cc"), "-o", tmp]
        subprocess.check_call(cmd)
        os.replace(tmp, LIB_PATH)                        # a reader never sees a half-written library
    return LIB_PATH


_L = None


def lib():
    global _L
    if _L is None:
        from ev2gym_b200 import _lib
        build()
        L = C.CDLL(LIB_PATH)
        L.ev2b_last_error.restype = C.c_char_p
        L.ev2b_last_error.argtypes = [C.c_void_p]
        L.ev2b_create.argtypes = [C.POINTER(_lib.Dims), C.POINTER(_lib.TopologyView), C.c_int, C.POINTER(C.c_void_p)]
        L.ev2b_destroy.restype = None
        L.ev2b_destroy.argtypes = [C.c_void_p]
        L.ev2b_obs_dim.argtypes = [C.c_void_p]
        L.ev2b_load_scenarios.argtypes = [C.c_void_p, C.POINTER(_lib.ScenariosView)]
        L.ev2b_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, _lib._pi, C.c_void_p, C.c_void_p]
        L.ev2b_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_lib.StepOut), C.c_void_p]
        L.ev2b_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ev2b_reset_done.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ev2b_state_view_get.argtypes = [C.c_void_p, C.POINTER(_lib.StateView)]
        L.ev2b_agent_actions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ev2b_step_k.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_int,
                                  C.POINTER(_lib.StepOut), C.c_void_p]
        L.ev2b_launch_count.restype = C.c_int64
        L.ev2b_launch_count.argtypes = [C.c_void_p]
        L.ev2b_kernel_launches.restype = C.c_int64
        L.ev2b_kernel_launches.argtypes = [C.c_void_p, C.c_int]
        L.ev2b_n_scenarios.argtypes = [C.c_void_p]
        _lib.declare_spawn(L)
        _L = L
    return _L


_OUT = {"reward": (np.float64, ()), "status": (np.uint32, ()), "obs": (np.float32, ("D",)),
        "cs_power": (np.float32, ("C",)), "cs_current": (np.float32, ("C",)), "tr_power": (np.float64, ("Tr",)),
        "tr_overload": (np.float64, ("Tr",)), "total_costs": (np.float64, ()), "action_mask": (np.uint8, ("P",)),
        "dep_sat": (np.float64, ("P",)), "dep_cap": (np.float64, ("P",)), "port_energy": (np.float32, ("P",)),
        "node_voltage": (np.float64, ("N",)), "hist_cs_power": (np.float32, ("T", "C")),
        "hist_cs_current": (np.float32, ("T", "C")), "hist_tr_overload": (np.float64, ("T", "Tr")),
        "hist_usage": (np.float64, ("T",))}


def _spawn_mixin():
    from ev2gym_b200.engine import SpawnMixin
    return SpawnMixin


class EmuEngine(_spawn_mixin()):
    """Same call sequence as ev2gym_b200.engine.BatchedEngine, numpy arrays instead of cuda tensors."""

    def __init__(self, topo, n_envs: int, reward=None, state=None, outputs: Iterable[str] = ("reward", "status", "obs"),
                 stats: bool = False):
        from ev2gym_b200 import _lib
        from ev2gym_b200.engine import REWARD_KINDS, STATE_KINDS, topology_view
        self._lib = _lib
        self.L = lib()
        self.topo, self.E = topo, int(n_envs)
        d = _lib.Dims(self.E, topo.C, topo.Tr, topo.T, topo.timescale, topo.dr_steps_ahead, REWARD_KINDS[reward],
                      STATE_KINDS[state], float(topo.tr_voltage), 1 if stats else 0, 0)
        tv, self._keep = topology_view(topo)
        h = C.c_void_p()
        rc = self.L.ev2b_create(C.byref(d), C.byref(tv), 0, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ev2b_create failed ({rc}): {self.L.ev2b_last_error(None).decode()}")
        self.h = h
        self.P, self.C_, self.Tr, self.T = topo.P, topo.C, topo.Tr, topo.T
        self.D = self.L.ev2b_obs_dim(self.h)
        self.set_outputs(outputs)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.L.ev2b_last_error(self.h).decode()}")

    def _stream(self):
        return None

    def close(self):
        if getattr(self, "h", None):
            self.L.ev2b_destroy(self.h)
            self.h = None

    def set_outputs(self, names):
        dims = {"D": max(self.D, 1), "C": self.C_, "Tr": self.Tr, "P": self.P, "N": self.topo.n_bus + 1, "T": self.T}
        self.out: Dict[str, np.ndarray] = {}
        self._so = self._lib.StepOut()
        for n in names:
            if n == "obs" and self.D == 0:
                continue
            dt, shp = _OUT[n]
            a = np.zeros((self.E,) + tuple(dims[k] for k in shp), dtype=dt)
            self.out[n] = a
            setattr(self._so, n, a.ctypes.data)

    def load_scenarios(self, scenarios):
        from ev2gym_b200.engine import scenarios_view
        v, _keep = scenarios_view(self.topo, scenarios)
        self._check(self.L.ev2b_load_scenarios(self.h, C.byref(v)), "ev2b_load_scenarios")

    def reset(self, env_lo=0, env_hi=None, scn_ids: Optional[Sequence[int]] = None):
        env_hi = self.E if env_hi is None else env_hi
        ids = None
        if scn_ids is not None:
            ids_np = np.ascontiguousarray(scn_ids, dtype=np.int32)
            ids = ids_np.ctypes.data_as(self._lib._pi)
        obs = self.out.get("obs")
        self._check(self.L.ev2b_reset(self.h, env_lo, env_hi, ids, obs.ctypes.data if obs is not None else None, None),
                    "ev2b_reset")
        return obs

    def reset_done(self):
        obs = self.out.get("obs")
        self._check(self.L.ev2b_reset_done(self.h, obs.ctypes.data if obs is not None else None, None), "ev2b_reset_done")
        return obs

    def step(self, actions: np.ndarray):
        assert actions.shape == (self.E, self.P) and actions.flags.c_contiguous
        dt = {np.dtype("float32"): 0, np.dtype("float64"): 1}[actions.dtype]
        self._check(self.L.ev2b_step(self.h, actions.ctypes.data, dt, C.byref(self._so), None), "ev2b_step")
        return self.out

    def step_host(self, actions, reward, status, obs=None):
        dt = {np.dtype("float32"): 0, np.dtype("float64"): 1}[actions.dtype]
        self._check(self.L.ev2b_step_host(self.h, actions.ctypes.data, dt, reward.ctypes.data, status.ctypes.data,
                                          obs.ctypes.data if obs is not None else None, None), "ev2b_step_host")

    AGENTS = {"external": 0, "afap": 1, "zero": 2, "uniform": 3, "roundrobin": 4, "calap": 5}

    def step_k(self, k, agent="afap", actions_k=None, seed=0, auto_reset=False):
        kind = self.AGENTS[agent]
        dt, ptr = 0, None
        if kind == 0:
            dt = 1 if actions_k.dtype == np.float64 else 0
            ptr = actions_k.ctypes.data
        low = -1.0 if self.topo.v2g_enabled else 0.0
        self._check(self.L.ev2b_step_k(self.h, int(k), kind, ptr, dt, int(seed), low, int(auto_reset),
                                       C.byref(self._so), None), "ev2b_step_k")
        return self.out

    def episode_stats(self):
        """get_statistics(env) of every env (needs stats=True), like BatchedEngine.episode_stats."""
        from ev2gym_b200.engine import STAT_NAMES
        out = np.zeros((self.E, len(STAT_NAMES)), dtype=np.float64)
        self.L.ev2b_episode_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.L.ev2b_episode_stats(self.h, out.ctypes.data, None), "ev2b_episode_stats")
        return {n: out[:, i].copy() for i, n in enumerate(STAT_NAMES)}

    def kernel_launches(self):
        """(step_kernel, evl_step_kernel, evl_rebuild_kernel) launches of this handle."""
        return tuple(int(self.L.ev2b_kernel_launches(self.h, k)) for k in range(3))

    def state(self) -> Dict[str, np.ndarray]:
        sv = self._lib.StateView()
        self._check(self.L.ev2b_state_view_get(self.h, C.byref(sv)), "ev2b_state_view_get")
        E, P = self.E, self.P

        def view(ptr, shape, dt):
            n = int(np.prod(shape))
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt).reshape(shape)
        return {"port_cap": view(sv.port_cap, (E, P), np.float64), "port_exch": view(sv.port_exch, (E, P), np.float64),
                "port_hot": view(sv.port_hot, (E, P, 4), np.int32), "env_step": view(sv.env_step, (E,), np.int32),
                "env_scn": view(sv.env_scn, (E,), np.int32), "env_potential": view(sv.env_potential, (E,), np.float64),
                "env_usage": view(sv.env_usage, (E,), np.float64), "env_kpi": view(sv.env_kpi, (E, sv.n_kpi), np.float64)}


class EmuTorchEngine(EmuEngine):
    """EmuEngine behind BatchedEngine's torch-facing surface (CPU tensors), so that the facades of ev2gym_b200.env can be
    driven without a GPU: tests monkeypatch `ev2gym_b200.env._ENGINE_CLS` with this class.  Test infrastructure only."""

    def __init__(self, topo, n_envs, reward=None, state=None, device=0, outputs=("reward", "status", "obs"), stats=False):
        import torch
        from ev2gym_b200.engine import _fn_name
        self.torch = torch
        self.dev = torch.device("cpu")
        self.reward_name, self.state_name = _fn_name(reward), _fn_name(state)
        super().__init__(topo, n_envs, reward=self.reward_name, state=self.state_name, outputs=outputs, stats=stats)

    def _t(self, d):
        return {k: self.torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v) for k, v in d.items()}

    def reset(self, env_lo=0, env_hi=None, scn_ids=None):
        obs = super().reset(env_lo, env_hi, scn_ids)
        return None if obs is None else self.torch.from_numpy(obs)

    def reset_done(self):
        obs = super().reset_done()
        return None if obs is None else self.torch.from_numpy(obs)

    def step(self, actions):
        a = np.ascontiguousarray(actions.detach().cpu().numpy())
        return self._t(super().step(a))

    def state_tensors(self):
        return self._t(self.state())

    @property
    def launch_count(self):
        return int(self.L.ev2b_launch_count(self.h))
