"""CPU-side checks of the CUDA kernels' LOGIC: ev2gym_b200/csrc compiled by g++ against the SIMT emulator in
tests/simt_emu (every CUDA thread a fiber; barriers, warp collectives, deferred cp.async, guard pages) and driven
through the same C ABI as libev2b.so, against the C oracle and the recorded reference traces.

This is test infrastructure: it proves indexing / barrier placement / list maintenance of the device code without a
GPU.  The parity tests proper are the `-m gpu` ones (tests/test_gpu_*.py); nothing in the product imports this.

Bars: battery level, list contents, counts, flags: bit-exact; float64 outputs 1e-9 relative; float32 outputs 1e-5.
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_emu"))

OUT = ("reward", "status", "obs", "tr_power", "tr_overload", "cs_power", "cs_current", "total_costs", "action_mask")


@pytest.fixture(scope="module")
def emu():
    import emu_engine
    emu_engine.build()
    return emu_engine


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def _occupied_ports(eng, e):
    st = eng.state()
    from ev2gym_b200.engine import BatchedEngine
    hot = BatchedEngine.decode_hot(st["port_hot"][e])
    t = int(st["env_step"][e])
    return np.nonzero((hot["t_arr"] <= t) & (t <= hot["t_dep"]))[0]


def _run_vs_oracle(emu, topo, bank, E, reward, state, adt, kernel, G=None, outputs=OUT, steps=None, switch=None,
                   monkeypatch=None):
    """Steps an emulated engine and the oracle side by side; `switch(t)` -> True asks for a per-port output on step t
    (under EV2B_EVL_MIX=1, a test-only knob, such a launch takes step_kernel and the list must be rebuilt afterwards)."""
    from oracle.oracle import OracleBatch
    monkeypatch.setenv("EV2B_KERNEL", kernel)
    if G:
        monkeypatch.setenv("EV2B_EVL_G", str(G))
    eng = emu.EmuEngine(topo, E, reward=reward, state=state, outputs=outputs)
    eng.load_scenarios(bank)
    scn_ids = [(3 * e + 1) % len(bank) for e in range(E)]
    obs0 = eng.reset(scn_ids=scn_ids)
    orc = OracleBatch(topo, [bank[i] for i in scn_ids], reward=reward, state=state)
    o0 = orc.reset()
    if eng.D:
        assert _close(obs0, o0, 1e-5, 1e-5)
    caps = eng.state()["port_cap"]
    rng = np.random.default_rng(99)
    Tr = topo.Tr
    n_invalid, n_served = np.zeros(E), np.zeros(E)
    for t in range(steps or topo.T):
        a = rng.uniform(-1, 1, (E, topo.P))
        a[rng.random((E, topo.P)) < 0.1] = 0.0
        if t % 9 == 4:
            a[:] = 1.0                      # everybody saturates: sum > 1 on shared chargers
        a = np.ascontiguousarray(a.astype(adt))
        if switch is not None:
            eng.set_outputs(outputs + (("port_energy",) if switch(t) else ()))
        out = eng.step(a)
        orc.step(a.astype(np.float64))
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(caps[occ], orc.arr["port_cap"][occ]), (t, "battery level")       # bit exact
        assert _close(out["reward"], orc.reward, 1e-9, 1e-9), (t, "reward", out["reward"], orc.reward)
        assert _close(out["total_costs"], [o.total_costs for o in orc.outs], 1e-9, 1e-12), (t, "costs")
        assert _close(out["tr_power"], orc.o["tr_power"][:, :Tr], 1e-9, 1e-9), (t, "tr_power")
        assert _close(out["tr_overload"], orc.o["tr_overload"][:, :Tr], 1e-9, 1e-9), (t, "tr_overload")
        assert _close(out["cs_power"], orc.o["cs_power"], 1e-5, 1e-6), (t, "cs_power")
        assert _close(out["cs_current"], orc.o["cs_current"], 1e-5, 1e-6), (t, "cs_current")
        if eng.D and (switch is None):      # (a fresh obs tensor every step would force full rewrites: checked separately)
            assert _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), (t, "obs")
        assert np.array_equal((out["status"] & 1) > 0, orc.done > 0), (t, "done")
        if "action_mask" in out:
            assert np.array_equal(out["action_mask"] > 0, occ), (t, "action mask")
        ovf = np.array([o.error == 1 for o in orc.outs])
        assert np.array_equal((out["status"] & 2) > 0, ovf), (t, "amps overflow flag")
        n_invalid += [o.invalid_actions for o in orc.outs]
        n_served += [o.n_departed for o in orc.outs]
        for e in (0, E - 1):                # the engine's own occupancy (hot words) agrees with the oracle's
            if not orc.done[e]:
                assert np.array_equal(_occupied_ports(eng, e), np.nonzero(occ[e])[0]), (t, e, "occupancy")
    kpi = eng.state()["env_kpi"]
    from ev2gym_b200.engine import KPI_NAMES
    k = {n: kpi[:, i] for i, n in enumerate(KPI_NAMES)}
    assert _close(k["total_reward"], [s.total_reward for s in orc.states], 1e-9, 1e-9)
    assert np.array_equal(k["total_evs_spawned"], [float(s.total_evs_spawned) for s in orc.states])
    assert np.array_equal(k["invalid_actions"], n_invalid) and np.array_equal(k["total_ev_served"], n_served)
    return eng, orc


SHAPES = [  # C, n_ports, Tr, E, reward, state, action dtype
    (25, 1, 1, 7, "SquaredTrackingErrorReward", "PublicPST", "float32"),
    (40, 2, 5, 5, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float32"),
    (150, 1, 1, 3, "profit_maximization", "V2G_profit_max", "float64"),
    (7, 3, 2, 9, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float64"),
    (40, 2, 5, 5, "V2G_profitmaxV2", "V2G_profit_max_loads", "float32"),
    (25, 1, 1, 6, "SqTrError_TrPenalty_UserIncentives", "PublicPST", "float64"),
    (30, 2, 3, 5, "SquaredTrackingErrorRewardWithPenalty", "PublicPST", "float32"),
    (30, 1, 2, 5, "pst_V2G_profitmaxV2", "V2G_profit_max", "float32"),
]


def _bank(C, n, Tr, T=48, n_scn=5):
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=C, n_ports=n, Tr=Tr, T=T, imin=6.0 if n == 3 else 0.0)
    return topo, sample_bank(topo, n_scn, seed=C + n, min_stay=5)


def test_emulator_runs_step_kernel_like_the_gpu(emu, monkeypatch):
    """The emulator itself: step_kernel (GPU-verified) must reproduce the oracle under emulation too."""
    topo, bank = _bank(40, 2, 5)
    eng, _ = _run_vs_oracle(emu, topo, bank, 5, SHAPES[1][4], SHAPES[1][5], "float32", "percharger", monkeypatch=monkeypatch)
    assert eng.kernel_launches() == (topo.T, 0, 0)
    eng.close()


@pytest.mark.parametrize("G", [1, 2, 4])
@pytest.mark.parametrize("C,n,Tr,E,reward,state,adt", SHAPES)
def test_evlist_kernel_matches_oracle(emu, C, n, Tr, E, reward, state, adt, G, monkeypatch):
    topo, bank = _bank(C, n, Tr)
    eng, _ = _run_vs_oracle(emu, topo, bank, E, reward, state, adt, "evlist", G=G, monkeypatch=monkeypatch)
    assert eng.kernel_launches() == (0, topo.T, 0)          # every launch took the event-driven kernel
    eng.close()


@pytest.mark.parametrize("G", [1, 2, 4])
def test_evlist_random_thread_schedule(emu, G, monkeypatch):
    """Same run with the emulator resuming threads in a seeded random order at every barrier round (a thread that runs
    ahead of its group must not see state another thread of the group is about to change)."""
    monkeypatch.setenv("SIMT_EMU_SEED", str(4 + G))
    topo, bank = _bank(40, 2, 5)
    eng, _ = _run_vs_oracle(emu, topo, bank, 5, SHAPES[1][4], SHAPES[1][5], "float32", "evlist", G=G, monkeypatch=monkeypatch)
    eng.close()


@pytest.mark.parametrize("G", [1, 4])
def test_evlist_list_is_the_set_of_connected_ports(emu, G, monkeypatch):
    """occ_list / occ_n after every step == the ports whose hot words say an EV is connected, in ascending port order
    (neighbouring threads of the EV loop then touch neighbouring ports: shared memory sectors)."""
    import ctypes as C
    from ev2gym_b200.engine import BatchedEngine
    topo, bank = _bank(40, 2, 5)
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    monkeypatch.setenv("EV2B_EVL_G", str(G))
    E = 4
    eng = emu.EmuEngine(topo, E, reward="profit_maximization", state="V2G_profit_max", outputs=("reward", "status"))
    eng.load_scenarios(bank)
    eng.reset()
    rng = np.random.default_rng(3)
    L = eng.L
    L.ev2b_debug_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ev2b_debug_list.restype = C.c_int
    for t in range(topo.T - 1):
        eng.step(np.ascontiguousarray(rng.uniform(-1, 1, (E, topo.P)).astype(np.float32)))
        for e in range(E):
            buf = np.zeros(topo.P, dtype=np.uint16)
            n = L.ev2b_debug_list(eng.h, e, buf.ctypes.data)
            got = [int(x) for x in buf[:n]]
            want = set(int(x) for x in _occupied_ports(eng, e))
            assert set(got) == want and len(got) == len(want), (t, e)
            assert got == sorted(got), (t, e, "port order")
    eng.close()


@pytest.mark.parametrize("n_ports", [1, 2, 3])
def test_evlist_list_in_port_order_for_every_charger_shape(emu, n_ports, monkeypatch):
    """The port-ordered list with one port per charger (no per-port staging, the flags exist only for the list), two (the
    paired CS phase) and three (the generic path)."""
    import ctypes as C
    topo, bank = _bank(24, n_ports, 3)
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    monkeypatch.setenv("EV2B_EVL_G", "1")
    eng = emu.EmuEngine(topo, 3, reward="profit_maximization", state="V2G_profit_max", outputs=("reward", "status"))
    eng.load_scenarios(bank)
    eng.reset()
    rng = np.random.default_rng(5)
    L = eng.L
    L.ev2b_debug_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ev2b_debug_list.restype = C.c_int
    for t in range(topo.T - 1):
        eng.step(np.ascontiguousarray(rng.uniform(-1, 1, (3, topo.P)).astype(np.float32)))
        for e in range(3):
            buf = np.zeros(topo.P, dtype=np.uint16)
            n = L.ev2b_debug_list(eng.h, e, buf.ctypes.data)
            got = [int(x) for x in buf[:n]]
            assert got == sorted(int(x) for x in _occupied_ports(eng, e)), (t, e)
    eng.close()


@pytest.mark.parametrize("adt", ["float32", "float64"])
def test_evlist_unaligned_action_buffer_takes_the_scalar_loads(emu, adt, monkeypatch):
    """Two ports per charger: both actions of a charger are read with ONE load when the caller's buffer sits on a
    two-element boundary.  A buffer that does not (a view one element into an allocation) must give the same episode."""
    topo, bank = _bank(20, 2, 2, T=30)
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    monkeypatch.setenv("EV2B_EVL_G", "1")
    E = 3
    engs = [emu.EmuEngine(topo, E, reward="profit_maximization", state="V2G_profit_max", outputs=("reward", "status", "obs"))
            for _ in range(2)]
    for g in engs:
        g.load_scenarios(bank)
        g.reset()
    rng = np.random.default_rng(11)
    item = np.dtype(adt).itemsize
    for t in range(topo.T):
        a = rng.uniform(-1, 1, (E, topo.P)).astype(adt)
        raw = np.zeros(E * topo.P + 4, dtype=adt)
        off = next(k for k in range(4) if (raw[k:].ctypes.data % (2 * item)) != 0)      # an odd element boundary
        shifted = raw[off:off + E * topo.P].reshape(E, topo.P)
        shifted[:] = a
        assert shifted.flags.c_contiguous and shifted.ctypes.data % (2 * item) != 0
        o0 = engs[0].step(np.ascontiguousarray(a))
        o1 = engs[1].step(shifted)
        for k in ("reward", "status", "obs"):
            assert np.array_equal(np.asarray(o0[k]), np.asarray(o1[k])), (t, k)
    assert np.array_equal(engs[0].state()["port_cap"], engs[1].state()["port_cap"])
    for g in engs:
        g.close()


@pytest.mark.parametrize("G", [1, 2])
def test_evlist_action_mask_incremental(emu, G, monkeypatch):
    """action_mask from the event-driven kernel: rows are updated in place (arrivals / departures only), rewritten when
    the caller hands in another buffer or an env starts a new episode (device-side auto reset)."""
    topo, bank = _bank(40, 2, 5, T=24)
    eng, orc = _run_vs_oracle(emu, topo, bank, 5, SHAPES[1][4], SHAPES[1][5], "float32", "evlist", G=G,
                              outputs=OUT, monkeypatch=monkeypatch)
    assert eng.kernel_launches() == (0, topo.T, 0)
    rng = np.random.default_rng(8)
    eng.out["action_mask"][:] = 1                    # stale terminal rows: the first step of the next episode must clear them
    eng.reset_done()
    for t in range(topo.T + 6):                      # through a whole episode and into the next one
        if t == 7:
            eng.set_outputs(OUT)                     # a fresh buffer mid-episode
            eng.out["action_mask"][:] = 1
        out = eng.step(np.ascontiguousarray(rng.uniform(-1, 1, (5, topo.P)).astype(np.float32)))
        for e in range(5):
            if not (out["status"][e] & 4):           # (a finished env keeps its last row, like step_kernel)
                want = np.zeros(topo.P, dtype=bool)
                if not (out["status"][e] & 1):
                    want[_occupied_ports(eng, e)] = True
                assert np.array_equal(out["action_mask"][e] > 0, want), (t, e)
        if t % 5 == 4:
            eng.reset_done()
    assert eng.kernel_launches()[0] == 0
    eng.close()


def test_evlist_and_step_kernel_interleave(emu, monkeypatch):
    """Both kernels on one handle (EV2B_EVL_MIX=1: launches that ask for port_energy take step_kernel); the next
    event-driven launch re-derives the list from the hot words."""
    monkeypatch.setenv("EV2B_EVL_MIX", "1")
    topo, bank = _bank(40, 2, 5)
    eng, _ = _run_vs_oracle(emu, topo, bank, 5, SHAPES[1][4], SHAPES[1][5], "float32", "evlist", G=2,
                            switch=lambda t: t % 5 in (1, 2), monkeypatch=monkeypatch)
    a, b, c = eng.kernel_launches()
    assert a > 0 and b > 0 and c > 0 and a + b == topo.T
    eng.close()


@pytest.mark.parametrize("kernel", ["percharger", "evlist"])
def test_emulated_kernels_agree_bitwise_on_state(emu, kernel, monkeypatch):
    """Both kernels leave identical hot / cap / exch arrays (they are interchangeable launch by launch)."""
    topo, bank = _bank(30, 2, 3)
    states = {}
    for kn in ("percharger", kernel):
        monkeypatch.setenv("EV2B_KERNEL", kn)
        eng = emu.EmuEngine(topo, 4, reward="ProfitMax_TrPenalty_UserIncentives", state="V2G_profit_max_loads")
        eng.load_scenarios(bank)
        eng.reset()
        rng = np.random.default_rng(1)
        obs = []
        for t in range(topo.T):
            out = eng.step(np.ascontiguousarray(rng.uniform(-1, 1, (4, topo.P)).astype(np.float32)))
            obs.append(out["obs"].copy())
        st = eng.state()
        states[kn] = (st["port_hot"].copy(), st["port_cap"].copy(), st["port_exch"].copy(), np.stack(obs),
                      st["env_kpi"].copy())
        eng.close()
    a, b = states["percharger"], states[kernel]
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    assert _close(a[4], b[4], 1e-12, 1e-12)


def test_evlist_auto_reset_step_k_and_host_chunks(emu, monkeypatch):
    """ev2b_step_k with the on-device UNIFORM agent + auto reset, and ev2b_step_host (two env chunks), on the
    event-driven kernel, against the same calls on step_kernel."""
    topo, bank = _bank(30, 2, 3, T=24)
    res = {}
    for kn in ("percharger", "evlist"):
        monkeypatch.setenv("EV2B_KERNEL", kn)
        monkeypatch.setenv("EV2B_EVL_G", "2")
        eng = emu.EmuEngine(topo, 6, reward="ProfitMax_TrPenalty_UserIncentives", state="V2G_profit_max_loads")
        eng.load_scenarios(bank)
        eng.reset()
        eng.step_k(40, agent="uniform", seed=11, auto_reset=True)        # crosses an episode boundary (T = 24)
        eng.step_k(9, agent="roundrobin", auto_reset=True)               # agent_kernel reads hot / cap, then a float64 step
        eng.step_k(9, agent="calap", auto_reset=True)
        eng.step_k(5, agent="afap", auto_reset=True)
        st = eng.state()
        snap = [st["port_hot"].copy(), st["port_cap"].copy(), st["env_step"].copy(), st["env_scn"].copy(),
                eng.out["obs"].copy()]
        rng = np.random.default_rng(4)
        rew, stat, obs = np.zeros(6), np.zeros(6, dtype=np.uint32), np.zeros((6, eng.D), dtype=np.float32)
        for t in range(5):
            eng.step_host(np.ascontiguousarray(rng.uniform(-1, 1, (6, topo.P))), rew, stat, obs)
        snap += [st["port_cap"].copy(), rew.copy(), stat.copy(), obs.copy()]
        res[kn] = snap
        if kn == "evlist":
            assert eng.kernel_launches()[0] == 0
        eng.close()
    for i, (x, y) in enumerate(zip(res["percharger"], res["evlist"])):
        if x.dtype.kind == "f" and i in (4, 6, 8):
            assert _close(x, y, 1e-6, 1e-6), i
        else:
            assert np.array_equal(x, y), i


def test_evlist_ragged_chargers(emu, monkeypatch):
    """Chargers with different port counts and ratings (topology-JSON case): the generic instantiation."""
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=12, n_ports=2, Tr=3, T=40)
    topo.cs_n_ports[:] = [1, 3, 2, 2, 4, 1, 2, 3, 1, 2, 2, 1]
    topo.cs_imax[::2] = 16.0
    bank = sample_bank(topo, 4, seed=5, min_stay=5)
    eng, _ = _run_vs_oracle(emu, topo, bank, 5, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float64",
                            "evlist", G=1, monkeypatch=monkeypatch)
    assert eng.kernel_launches() == (0, topo.T, 0)
    eng.close()


def _lean_golden():
    out = []
    for name in golden_cases():
        tr = np.load(f"{GOLDEN}/{name}.trace.npz")
        if not name.startswith("grid"):          # every stock reward / state function that needs no distribution grid
            out.append(name)
    return out


@pytest.mark.parametrize("name", _lean_golden())
def test_evlist_kernel_matches_reference_trace(emu, name, monkeypatch):
    """The recorded episodes of the Python reference (tests/golden) through the event-driven kernel."""
    from ev2gym_b200.scenario import ScenarioPack
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    topo = pack.topo
    E = 3
    eng = emu.EmuEngine(topo, E, reward=str(tr["reward_fn"]), state=str(tr["state_fn"]), outputs=OUT)
    eng.load_scenarios(pack.scenarios)
    obs0 = eng.reset()
    assert _close(obs0[0], tr["obs0"], 1e-5, 1e-6)
    st = eng.state()
    T = tr["reward"].shape[0]
    for t in range(T):
        out = eng.step(np.ascontiguousarray(np.tile(tr["actions"][t], (E, 1)).astype(np.float64)))
        for e in (0, E - 1):
            occ = tr["action_mask"][t] > 0
            assert np.array_equal(_occupied_ports(eng, e), np.nonzero(occ)[0]) or bool(tr["done"][t]), (t, "occupancy")
            assert np.array_equal(st["port_cap"][e][occ], tr["cap"][t][occ]), (t, "cap")
            assert np.array_equal(out["action_mask"][e] > 0, occ), (t, "action mask")
            assert _close(out["reward"][e], tr["reward"][t], 1e-9, 1e-9), (t, out["reward"][e], tr["reward"][t])
            assert _close(out["total_costs"][e], tr["total_costs"][t], 1e-9, 1e-12)
            assert _close(out["tr_power"][e], tr["tr_power"][t], 1e-9, 1e-9)
            assert _close(out["tr_overload"][e], tr["tr_overload"][t], 1e-9, 1e-9)
            assert _close(out["cs_power"][e], tr["cs_power"][t], 1e-5, 1e-6)
            assert _close(out["cs_current"][e], tr["cs_current"][t], 1e-5, 1e-6)
            assert _close(out["obs"][e], tr["obs"][t], 1e-5, 1e-5), (t, "obs")
            assert bool(out["status"][e] & 1) == bool(tr["done"][t])
    from ev2gym_b200.engine import KPI_NAMES
    k = {n: st["env_kpi"][:, i] for i, n in enumerate(KPI_NAMES)}
    assert k["total_reward"][0] == pytest.approx(float(tr["total_reward"]), rel=1e-9, abs=1e-9)
    assert k["total_ev_served"][0] == float(tr["stat_total_ev_served"])
    assert k["total_energy_charged"][0] == pytest.approx(float(tr["stat_total_energy_charged"]), rel=1e-9)
    assert eng.kernel_launches() == (0, T, 0)
    out = eng.step(np.zeros((E, topo.P)))      # stepping a finished env: WAS_DONE, reward 0 (ev2gym_env.py:343)
    assert int(out["status"][0]) & 4 and float(out["reward"][0]) == 0.0
    eng.close()


BENCH_PACKS = [("c2_publicpst_c25", "SquaredTrackingErrorReward", "PublicPST", 1),
               ("c3_v2gloads_c100n2tr5", "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", 2),
               ("c4_v2gprofitmax_c250", "profit_maximization", "V2G_profit_max", 2)]


@pytest.mark.parametrize("pack_name,reward,state,G", BENCH_PACKS)
def test_evlist_on_bench_scenario_banks(emu, pack_name, reward, state, G, monkeypatch):
    """tests/test_gpu_fullsize.py on the emulator, on a sample of the reference-exported banks bench.py uses (16 scenarios
    x 2 replicas, whole episodes, 70 % of the ports occupied at the busy part): replicas agree bitwise, battery levels
    and action mask equal the oracle's, reward 1e-9, observation 1e-5."""
    from ev2gym_b200.scenario import ScenarioPack
    from oracle.oracle import OracleBatch
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    monkeypatch.setenv("EV2B_EVL_G", str(G))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pack = ScenarioPack.load(os.path.join(root, "ev2gym_b200", "data", pack_name + ".npz"))
    topo, S = pack.topo, 16
    scen = pack.scenarios[:S]
    E = 2 * S
    eng = emu.EmuEngine(topo, E, reward=reward, state=state, outputs=("reward", "status", "obs", "action_mask"))
    eng.load_scenarios(scen)
    obs0 = eng.reset()
    orc = OracleBatch(topo, scen, reward=reward, state=state)
    assert _close(obs0[:S], orc.reset(), 1e-5, 1e-5)
    caps = eng.state()["port_cap"]
    low = -1.0 if topo.v2g_enabled else 0.0
    rng = np.random.default_rng(11)
    for step in range(topo.T):
        a = rng.uniform(low, 1.0, (S, topo.P))
        a[rng.random((S, topo.P)) < 0.1] = 0.0
        out = eng.step(np.ascontiguousarray(np.tile(a, (2, 1)).astype(np.float32)))
        for k in ("reward", "status", "obs", "action_mask"):
            v = out[k].reshape((2, S) + out[k].shape[1:])
            assert (v == v[:1]).all(), (step, k)
        orc.step(a.astype(np.float32).astype(np.float64))
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(caps[:S][occ], orc.arr["port_cap"][occ]), step
        assert np.array_equal(out["action_mask"][:S] > 0, occ), (step, "action mask")
        assert _close(out["reward"][:S], orc.reward, 1e-9, 1e-9), step
        assert _close(out["obs"][:S], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), step
        assert np.array_equal((out["status"][:S] & 1).astype(bool), orc.done.astype(bool)), step
    assert (out["status"] & 1).all() and eng.kernel_launches() == (0, topo.T, 0)
    eng.close()


ALL_OUT = ("reward", "status", "obs", "cs_power", "cs_current", "tr_power", "tr_overload", "total_costs", "action_mask",
           "dep_sat", "dep_cap", "port_energy")


@pytest.mark.parametrize("kernel", ["percharger", "evlist"])
@pytest.mark.parametrize("name", golden_cases())
def test_full_featured_kernels_match_reference_trace_on_emulator(emu, name, kernel, monkeypatch):
    """tests/test_gpu_parity.py::test_cuda_matches_reference_trace on the emulator: both step kernels in their
    full-featured instantiation (statistics mode, every output, the Laurent power flow of the grid episodes) on all
    recorded reference episodes, including get_statistics() at the end -- so a change to the shared model code
    (ev_step_item) is checked here before it reaches a GPU."""
    from ev2gym_b200.scenario import ScenarioPack
    monkeypatch.setenv("EV2B_KERNEL", kernel)
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    topo = pack.topo
    E, grid = 2, pack.topo.n_bus > 0
    eng = emu.EmuEngine(topo, E, reward=str(tr["reward_fn"]), state=str(tr["state_fn"]),
                        outputs=ALL_OUT + (("node_voltage",) if grid else ()), stats=True)
    eng.load_scenarios(pack.scenarios)
    obs0 = eng.reset()
    assert _close(obs0[0], tr["obs0"], 1e-5, 1e-6)
    st = eng.state()
    T = tr["reward"].shape[0]
    for t in range(T):
        out = eng.step(np.ascontiguousarray(np.tile(tr["actions"][t], (E, 1)).astype(np.float64)))
        e = E - 1
        occ = out["action_mask"][e] > 0
        assert np.array_equal(occ.astype(np.float64), tr["action_mask"][t]), (t, "mask")
        assert np.array_equal(st["port_cap"][e][occ], tr["cap"][t][occ]), (t, "cap")              # bit exact
        assert _close(out["reward"][e], tr["reward"][t], 1e-9, 1e-9), (t, "reward")
        assert _close(out["total_costs"][e], tr["total_costs"][t], 1e-9, 1e-12)
        assert _close(out["tr_power"][e], tr["tr_power"][t], 1e-9, 1e-9)
        assert _close(out["tr_overload"][e], tr["tr_overload"][t], 1e-9, 1e-9)
        assert _close(out["cs_power"][e], tr["cs_power"][t], 1e-5, 1e-6)
        assert _close(out["obs"][e], tr["obs"][t], 1e-5, 1e-5), (t, "obs")
        if grid:
            assert _close(out["node_voltage"][e], tr["node_voltage"][:, t], 1e-9, 1e-12), (t, "node voltage")
        assert bool(out["status"][e] & 1) == bool(tr["done"][t])
        assert np.count_nonzero(~np.isnan(out["dep_sat"][e])) == tr["n_departed"][t]
        assert np.array_equal(np.isnan(out["dep_sat"][e]), np.isnan(out["dep_cap"][e]))
        if "port_energy" in tr.files:
            assert _close(out["port_energy"][e], tr["port_energy"][t], 1e-5, 1e-6), (t, "port energy")
    assert eng.kernel_launches()[1 if kernel == "percharger" else 0] == 0
    stats = eng.episode_stats()
    from ev2gym_b200.engine import STAT_NAMES
    for n in STAT_NAMES:
        ref = float(tr["stat_" + n])
        got = float(stats[n][E - 1])
        assert (np.isnan(ref) and np.isnan(got)) or got == pytest.approx(ref, rel=1e-9, abs=1e-9), (n, got, ref)
    eng.close()


@pytest.mark.parametrize("G", [1, 2, 4])
@pytest.mark.parametrize("agent", ["uniform", "external", "afap"])
def test_kstep_kernel_equals_launch_per_step(emu, G, agent, monkeypatch):
    """ev2b_step_k on the event-driven kernel: ONE launch that advances every env k steps (evl_step_kernel<KSTEP>, the
    group loops over its env; device-side reset of finished envs) leaves exactly the state and outputs of k launches +
    ev2b_reset_done (EV2B_STEP_K=loop), and the external-action variant equals the oracle stepped with the same tensor."""
    from oracle.oracle import OracleBatch
    topo, bank = _bank(30, 2, 3, T=24)
    E = 7
    rng = np.random.default_rng(5)
    ks = (3, 1, 17, 9, 30)                                 # crosses two episode boundaries (T = 24)
    acts = [np.ascontiguousarray(rng.uniform(-1, 1, (k, E, topo.P)).astype(np.float32)) for k in ks]
    res = {}
    for mode in ("loop", "kstep"):
        monkeypatch.setenv("EV2B_KERNEL", "evlist")
        monkeypatch.setenv("EV2B_EVL_G", str(G))
        monkeypatch.setenv("EV2B_STEP_K", "loop" if mode == "loop" else "fused")
        eng = emu.EmuEngine(topo, E, reward="ProfitMax_TrPenalty_UserIncentives", state="V2G_profit_max_loads",
                            outputs=("reward", "status", "obs", "action_mask"))
        eng.load_scenarios(bank)
        eng.reset()
        snaps, n0 = [], eng.kernel_launches()[1]
        for k, a in zip(ks, acts):
            out = eng.step_k(k, agent=agent, actions_k=a if agent == "external" else None, seed=77, auto_reset=True)
            st = eng.state()
            snaps.append([st[n].copy() for n in ("port_hot", "port_cap", "port_exch", "env_step", "env_scn", "env_kpi")] +
                         [out[n].copy() for n in ("reward", "status", "obs", "action_mask")])
        n_launch = eng.kernel_launches()[1] - n0
        assert n_launch == (sum(ks) if mode == "loop" else sum(1 if k > 1 else 1 for k in ks)), (mode, n_launch)
        res[mode] = snaps
        eng.close()
    for a, b in zip(res["loop"], res["kstep"]):
        for i, (x, y) in enumerate(zip(a, b)):
            assert np.array_equal(x, y), i
    if agent == "external":                                # no auto reset here: the oracle has none
        monkeypatch.setenv("EV2B_STEP_K", "fused")
        eng = emu.EmuEngine(topo, E, reward="ProfitMax_TrPenalty_UserIncentives", state="V2G_profit_max_loads")
        eng.load_scenarios(bank)
        eng.reset()
        orc = OracleBatch(topo, [bank[e % len(bank)] for e in range(E)], reward="ProfitMax_TrPenalty_UserIncentives",
                          state="V2G_profit_max_loads")
        orc.reset()
        a = np.ascontiguousarray(rng.uniform(-1, 1, (topo.T, E, topo.P)))
        out = eng.step_k(topo.T, agent="external", actions_k=a)
        for t in range(topo.T):
            orc.step(a[t])
        assert _close(out["reward"], orc.reward, 1e-9, 1e-9) and _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5)
        assert (out["status"] & 1).all()
        kpi = eng.state()["env_kpi"]
        assert _close(kpi[:, 0], [s.total_reward for s in orc.states], 1e-9, 1e-9)
        eng.close()


@pytest.mark.parametrize("kernel", ["percharger", "evlist"])
def test_more_than_1023_ports_per_env(emu, kernel, monkeypatch):
    """600 chargers x 2 ports: the env-level counts (empty ports, departures, arrivals of a step) exceed the 10-bit fields
    step_kernel used to pack them in (round-1 advisor finding); the KPI sums must still equal the oracle's."""
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=600, n_ports=2, Tr=3, T=10)
    bank = sample_bank(topo, 2, seed=3, min_stay=3)
    eng, orc = _run_vs_oracle(emu, topo, bank, 2, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float64",
                              kernel, G=2, monkeypatch=monkeypatch)
    assert eng.state()["env_kpi"][:, 11].min() > 1023          # invalid_actions: more than one step's worth of 10 bits
    eng.close()


@pytest.mark.parametrize("kernel,n", [("percharger", 2), ("evlist", 2), ("evlist", 1)])
def test_history_outputs_equal_the_oracles_histories(emu, kernel, n, monkeypatch):
    """hist_cs_power / hist_cs_current / hist_tr_overload / hist_usage after an episode == what the reference keeps in
    env.cs_power[C,T], env.cs_current[C,T], env.tr_overload[Tr,T], env.current_power_usage[T] (utils.py:794-861,
    ev2gym_env.py:520-556; here: the oracle's restatement of them), for every env of the batch."""
    topo, bank = _bank(24, n, 3, T=30)
    hist = ("hist_cs_power", "hist_cs_current", "hist_tr_overload", "hist_usage")
    eng, orc = _run_vs_oracle(emu, topo, bank, 4, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float32",
                              kernel, G=2, outputs=OUT + hist, monkeypatch=monkeypatch)
    E, C, T, Tr = 4, topo.C, topo.T, topo.Tr
    assert _close(eng.out["hist_cs_power"].transpose(0, 2, 1), orc.arr["cs_power_hist"].reshape(E, C, T), 1e-5, 1e-6)
    assert _close(eng.out["hist_cs_current"].transpose(0, 2, 1), orc.arr["cs_current_hist"].reshape(E, C, T), 1e-5, 1e-6)
    assert _close(eng.out["hist_tr_overload"].transpose(0, 2, 1), orc.arr["tr_overload_hist"].reshape(E, -1, T)[:, :Tr], 1e-9, 1e-9)
    assert _close(eng.out["hist_usage"], orc.arr["usage"].reshape(E, -1)[:, :T], 1e-9, 1e-9)
    assert np.abs(eng.out["hist_cs_power"]).max() > 0
    eng.close()


@pytest.mark.parametrize("shape", [0, 1, 3])
def test_evlist_one_env_per_cta(emu, shape, monkeypatch):
    """The 32-thread instantiation (one warp = one env = one CTA; chosen for launches of more than one wave, forced here
    with EV2B_EVL_TPB=32) computes the same thing, incl. the k-step kernel with device-side reset."""
    C, n, Tr, E, reward, state, adt = SHAPES[shape]
    monkeypatch.setenv("EV2B_EVL_TPB", "32")
    topo, bank = _bank(C, n, Tr)
    eng, _ = _run_vs_oracle(emu, topo, bank, E, reward, state, adt, "evlist", G=1, monkeypatch=monkeypatch)
    assert eng.kernel_launches() == (0, topo.T, 0)
    eng.reset()
    eng.step_k(topo.T + 5, agent="uniform", seed=3, auto_reset=True)
    st = eng.state()
    assert (st["env_step"] == 5).all()
    eng.close()
