"""Device-side generate_power_setpoints (ev2gym_b200/csrc/ev2b_spawn.cuh, spawn_setpoints_kernel) against the REFERENCE's
own function on the same random numbers.

The device sampler draws EV sessions and then derives the power setpoints from them like ev2gym/utilities/utils.py:664-757.
Its random numbers come from a counter-based generator, the reference's from numpy's global stream, so tests/test_spawn.py
can only compare distributions.  Here the comparison is exact: the sessions the device drew are handed to the UNMODIFIED
reference function as `env.EVs_profiles`, and `np.random.normal` is replaced, for the duration of the call, by a function
that returns the device generator's normals for the EV being processed (splitmix64 counter + Box-Muller, restated below
from ev2b_spawn.cuh).  Window, weights, normalisation, the <= 11 repair sweeps, the sum and the median filter are then the
reference's code, and the result must equal what the kernel left in the bank (1e-9: summation order, libm).

CPU only (the kernel runs on the SIMT emulator, tests/simt_emu); skipped where /root/reference is absent.
"""
import math
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "ev2gym_b200", "data")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ev2gym")), reason="needs the reference checkout")

M64 = (1 << 64) - 1


def _uniform(seed, counter):
    """spawn_uniform (ev2b_spawn.cuh): splitmix64 of (seed, counter) -> [0, 1)."""
    z = (seed + (counter + 1) * 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return float(z >> 11) * (1.0 / 9007199254740992.0)


def _normal(seed, c, mean, sd):
    """spawn_normal: Box-Muller on two consecutive uniforms."""
    u1, u2 = 1.0 - _uniform(seed, c), _uniform(seed, c + 1)
    return mean + sd * math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586 * u2)


def test_device_setpoints_equal_the_reference_function_on_the_same_normals(monkeypatch):
    old_path = list(sys.path)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    try:
        from ev2gym.utilities import utils as ref_utils
        import emu_engine
        from ev2gym_b200.scenario import ScenarioPack, SpawnTables
        emu_engine.build()
        monkeypatch.setenv("EV2B_KERNEL", "evlist")
        name = "c2_publicpst_c25"
        pack = ScenarioPack.load(os.path.join(DATA, name + ".npz"))
        tab = SpawnTables.load(os.path.join(DATA, "spawn_" + name + ".npz"))
        assert tab.power_setpoint_enabled
        topo, S = pack.topo, 12
        T, P = topo.T, topo.P
        eng = emu_engine.EmuEngine(topo, S, reward="SquaredTrackingErrorReward", state="PublicPST", outputs=("reward",))
        eng.set_spawn_tables(tab)
        eng.load_scenarios(pack.scenarios[:S])
        seed = 0x1234ABCD5678
        eng.resample_sessions(seed=seed)
        # charging_stations[0].get_min_charge_power() / get_max_power()   ev_charger.py:251-255
        v, ph = float(topo.cs_voltage[0]), float(topo.cs_phases[0])
        min_cs = float(topo.cs_imin[0]) * v * math.sqrt(ph) / 1000.0
        max_cs = float(topo.cs_imax[0]) * v * math.sqrt(ph) / 1000.0
        cs0 = types.SimpleNamespace(get_min_charge_power=lambda: min_cs, get_max_power=lambda: max_cs)
        n_checked = 0
        for s in range(S):
            d = eng.read_sessions(s)
            m = d["model"]
            evs = [types.SimpleNamespace(time_of_arrival=int(ta), time_of_departure=int(td), battery_capacity=float(tab.model_B[k]),
                                         battery_capacity_at_arrival=float(c0), min_ac_charge_power=float(tab.model_pmin_ac[k]),
                                         max_ac_charge_power=float(tab.model_pmax_ac[k]))
                   for ta, td, k, c0 in zip(d["t_arr"], d["t_dep"], m, d["cap0"])]
            env = types.SimpleNamespace(simulation_length=T, timescale=topo.timescale,
                                        charge_prices=np.asarray(pack.scenarios[s].charge_price, dtype=np.float64)[None, :T],
                                        config={"power_setpoint_flexiblity": tab.power_setpoint_flexibility},
                                        charging_stations=[cs0], EVs_profiles=evs)
            calls = iter(zip(d["port"], d["t_arr"]))

            def device_normals(loc, scale, size):
                port, ta = next(calls)                    # the reference walks EVs_profiles in order: this call is that EV's
                c = (1 << 62) + ((s * P + int(port)) * T + int(ta)) * 2 * T
                loc = np.asarray(loc, dtype=np.float64)
                assert loc.shape == (size,)
                return np.array([_normal(seed, c + 2 * i, float(loc[i]), float(scale)) for i in range(size)])
            monkeypatch.setattr(np.random, "normal", device_normals)
            want = np.asarray(ref_utils.generate_power_setpoints(env), dtype=np.float64)
            monkeypatch.undo()
            monkeypatch.setenv("EV2B_KERNEL", "evlist")
            assert next(calls, None) is None, "the reference consumed one draw per EV"
            got = eng.read_setpoints(s)
            assert got.shape == want.shape
            assert np.all(np.abs(got - want) <= 1e-9 + 1e-9 * np.abs(want)), (s, float(np.abs(got - want).max()))
            n_checked += len(evs)
        assert n_checked > 100
        eng.close()
    finally:
        sys.path[:] = old_path
