#!/usr/bin/env python
"""Seconds-long GPU check without torch: drives libev2b.so through ctypes with cudaMalloc'ed buffers (libcudart via
ctypes) and compares a few short episodes of the event-driven kernel with the C oracle.  For GPU slots too short for
`pytest -m gpu` (a fresh box needs ~1 min for the first `import torch`).  Writes gpurun_out/raw_smoke.log as it goes."""
import ctypes as C
import os
import sys
import time

import numpy as np

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "raw_smoke.log"), "w")


def log(*a):
    print(f"[{time.time() - T0:6.2f}s]", *a, file=LOG, flush=True)
    print(f"[{time.time() - T0:6.2f}s]", *a, flush=True)


EMU = os.environ.get("RAW_SMOKE_EMU") == "1"      # dry run of this script against the SIMT-emulator build (no GPU)
if EMU:
    _libc = C.CDLL(None)
    _libc.malloc.restype, _libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    _libc.memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]

    class _FakeRt:
        @staticmethod
        def cudaMalloc(pp, n):
            pp._obj.value = _libc.malloc(n)
            return 0

        @staticmethod
        def cudaMemcpy(d, s, n, kind):
            _libc.memcpy(d, s, n)
            return 0
    rt = _FakeRt()
else:
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12" if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else "libcudart.so.12")
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]


class DevArray:
    def __init__(self, shape, dtype):
        self.host = np.zeros(shape, dtype=dtype)
        self.ptr = C.c_void_p()
        assert rt.cudaMalloc(C.byref(self.ptr), max(1, self.host.nbytes)) == 0
        self.up()

    def up(self):
        assert rt.cudaMemcpy(self.ptr, self.host.ctypes.data, self.host.nbytes, 1) == 0

    def down(self):
        assert rt.cudaMemcpy(self.host.ctypes.data, self.ptr, self.host.nbytes, 2) == 0
        return self.host


def run(L, _lib, topo, bank, E, reward, state, G, outputs):
    from ev2gym_b200.engine import REWARD_KINDS, STATE_KINDS, scenarios_view, topology_view
    from oracle.oracle import OracleBatch
    os.environ["EV2B_KERNEL"], os.environ["EV2B_EVL_G"] = "evlist", str(G)
    d = _lib.Dims(E, topo.C, topo.Tr, topo.T, topo.timescale, topo.dr_steps_ahead, REWARD_KINDS[reward], STATE_KINDS[state],
                  float(topo.tr_voltage), 0, 0)
    tv, keep = topology_view(topo)
    h = C.c_void_p()
    assert L.ev2b_create(C.byref(d), C.byref(tv), 0, C.byref(h)) == 0, L.ev2b_last_error(None)
    sv, keep2 = scenarios_view(topo, bank)
    assert L.ev2b_load_scenarios(h, C.byref(sv)) == 0, L.ev2b_last_error(h)
    D = L.ev2b_obs_dim(h)
    shapes = {"reward": ((E,), np.float64), "status": ((E,), np.uint32), "obs": ((E, max(D, 1)), np.float32),
              "action_mask": ((E, topo.P), np.uint8), "tr_power": ((E, topo.Tr), np.float64),
              "cs_power": ((E, topo.C), np.float32)}
    out = {k: DevArray(*shapes[k]) for k in outputs}
    so = _lib.StepOut()
    for k, v in out.items():
        setattr(so, k, v.ptr.value)
    assert L.ev2b_reset(h, 0, E, None, out["obs"].ptr, None) == 0
    st = _lib.StateView()
    assert L.ev2b_state_view_get(h, C.byref(st)) == 0
    caps = np.zeros((E, topo.P))
    orc = OracleBatch(topo, [bank[e % len(bank)] for e in range(E)], reward=reward, state=state)
    orc.reset()
    act = DevArray((E, topo.P), np.float32)
    rng = np.random.default_rng(0)
    for t in range(topo.T):
        act.host[:] = rng.uniform(-1, 1, (E, topo.P))
        if t % 9 == 4:
            act.host[:] = 1.0
        act.up()
        assert L.ev2b_step(h, act.ptr, 0, C.byref(so), None) == 0, L.ev2b_last_error(h)
        orc.step(act.host.astype(np.float64))
        assert rt.cudaMemcpy(caps.ctypes.data, C.c_void_p(st.port_cap), caps.nbytes, 2) == 0
        occ = orc.arr["port_session"] >= 0
        o = {k: v.down() for k, v in out.items()}
        assert np.array_equal(caps[occ], orc.arr["port_cap"][occ]), (t, "cap")
        assert np.allclose(o["reward"], orc.reward, rtol=1e-9, atol=1e-9), (t, "reward")
        assert np.allclose(o["obs"][:, :D], orc.o["obs"][:, :D], rtol=1e-5, atol=1e-5), (t, "obs")
        assert np.array_equal(o["action_mask"] > 0, occ), (t, "mask")
        assert np.allclose(o["tr_power"], orc.o["tr_power"][:, :topo.Tr], rtol=1e-9, atol=1e-9), (t, "tr_power")
        assert np.allclose(o["cs_power"], orc.o["cs_power"], rtol=1e-5, atol=1e-6), (t, "cs_power")
        assert np.array_equal((o["status"] & 1) > 0, orc.done > 0), (t, "done")
    by = tuple(int(L.ev2b_kernel_launches(h, k)) for k in range(3))
    assert by == (0, topo.T, 0), by
    L.ev2b_destroy(h)


def main():
    from ev2gym_b200 import _lib
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    log("imports done")
    if EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
        import emu_engine
        L = emu_engine.lib()
    else:
        L = _lib.load()
    L.ev2b_kernel_launches.restype = C.c_int64
    log("library loaded")
    outs = ("reward", "status", "obs", "action_mask", "tr_power", "cs_power")
    cases = [(40, 2, 5, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", 2),
             (60, 1, 1, "profit_maximization", "V2G_profit_max", 2),
             (25, 1, 1, "SquaredTrackingErrorReward", "PublicPST", 1),
             (40, 2, 3, "V2G_profitmaxV2", "V2G_profit_max_loads", 4)]
    for C_, n, Tr, rw, stt, G in cases:
        topo = Topology.uniform(C=C_, n_ports=n, Tr=Tr, T=40)
        bank = sample_bank(topo, 4, seed=C_ + n, min_stay=5)
        run(L, _lib, topo, bank, 48, rw, stt, G, outs)
        log("ok", C_, n, Tr, rw, stt, "G", G)
    log("ALL OK")


if __name__ == "__main__":
    try:
        main()
    except BaseException as exc:  # noqa: BLE001 -- the log must say what happened
        log("FAILED", repr(exc))
        raise
