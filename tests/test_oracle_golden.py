"""CPU: pin the C oracle (oracle/ev2o.c) against traces recorded from the unmodified Python
reference (tools/make_golden.py) and against the known-answer table of SURVEY.md section 8c.

Bar: BIT-EXACT float64 on every per-step quantity (the oracle restates the reference's
operation order literally and is compiled without FMA contraction)."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from ev2gym_b200.scenario import ScenarioPack, Topology, Scenario, assign_ports
from oracle.oracle import OracleEnv, lib

import ctypes as C


def _eq(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.array_equal(a[~both_nan], b[~both_nan]) and a.shape == b.shape


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_trace_bit_exact(name):
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    env = OracleEnv(pack.topo, pack.scenarios[0], reward=str(tr["reward_fn"]), state=str(tr["state_fn"]))
    assert _eq(env.reset(), tr["obs0"])
    grid = "node_voltage" in tr.files
    T = tr["reward"].shape[0]
    for t in range(T):
        r = env.step(tr["actions"][t])
        assert r["error"] == 0
        for k in ("obs", "cs_power", "cs_current", "tr_power", "tr_amps", "tr_overload", "action_mask",
                  "actions_eff"):
            assert _eq(r[k], tr[k][t]), (k, t)
        if grid:      # BLAS zgemv sums the 33 products of a power-flow row in its own order: 1e-12, not bit-exact
            assert r["reward"] == pytest.approx(tr["reward"][t], rel=1e-12, abs=1e-12), t
            assert np.allclose(r["node_vm"], tr["node_voltage"][:, t], rtol=0, atol=1e-13), t
        else:
            assert r["reward"] == tr["reward"][t], t
        assert r["total_costs"] == tr["total_costs"][t], t
        assert r["invalid_actions"] == tr["invalid"][t] and r["n_departed"] == tr["n_departed"][t], t
        occ = r["port_session"] >= 0
        assert _eq(np.where(occ, r["port_cap"], np.nan), tr["cap"][t]), t
        assert _eq(np.where(occ, r["port_energy_exch"], 0.0), tr["energy_exch"][t]), t
        # arrival / departure indexing: the EV sitting in each port is the same EV
        sess_arr = np.where(occ, pack.scenarios[0].sessions["t_arr"][np.maximum(r["port_session"], 0)], -1)
        assert np.array_equal(sess_arr, tr["port_t_arr"][t]), t
        assert r["done"] == bool(tr["done"][t])
    assert _eq(r["usage"], tr["usage"]) and _eq(r["potential"], tr["potential"])
    assert env.total_reward == pytest.approx(float(tr["total_reward"]), rel=1e-12 if grid else 0, abs=0)
    with pytest.raises(AssertionError):
        env.step(tr["actions"][0])      # ev2gym_env.py:343


@pytest.mark.parametrize("name", golden_cases())
def test_host_port_assignment_matches_reference(name):
    """`assign_ports` (host replay of evs_connected.index(None)) == where the reference put each EV."""
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    s = pack.scenarios[0].sessions
    port = assign_ports(pack.topo, s["t_arr"], s["t_dep"], s["loc"])
    T = tr["reward"].shape[0]
    for i in range(len(port)):
        ta = int(s["t_arr"][i])
        if ta - 1 < T:   # recorded after step ta-1: the EV that arrived at ta sits in port[i]
            assert tr["port_t_arr"][ta - 1][port[i]] == ta


def _ev_step(cap, amps, ts=1.0, mult=1.0, B=50.0, pmax=11.0, pdis=-11.0, bmin=5.0, eta=1.0, V=400.0, ph=3, dt=15):
    q = (C.c_double * 11)(cap, B, pmax, 0.0, pdis, 0.0, bmin, ts, mult, eta, eta)
    e, a = C.c_double(), C.c_double()
    new = lib().ev2o_ev_step(q, 3, amps, V, ph, dt, C.byref(e), C.byref(a))
    return new, e.value, a.value


def test_known_answers_ev_step():
    """SURVEY.md section 8c KA0-KA2, KA4 (values produced there by the by-path Python oracle)."""
    cap = 20.0
    seq = []
    for _ in range(3):
        cap, e, a = _ev_step(cap, 32.0)
        seq.append(cap)
    assert seq == [22.75, 25.5, 28.26]                       # KA0: the ceil lands on 28.26, not 28.25
    assert abs(a - 15.8771) < 1e-4 and abs(e * 4 - 11.0) < 1e-9
    cap, seq = 38.0, []
    for _ in range(4):
        cap, e, a = _ev_step(cap, 32.0, ts=0.8, mult=5.0)
        seq.append(cap)
    assert seq[:3] == [40.75, 43.5, 46.25]                   # KA1 (pre-transition steps)
    cap, e, a = _ev_step(38.0, 32.0, ts=0.8, mult=5.0)
    assert e == 2.7500000000000027 and a == 15.877132402714725
    cap, e, a = _ev_step(7.0, -32.0)                         # KA2
    assert cap == 5.0 and e == -2.0 and a == -11.547005383792516
    cap, e, a = _ev_step(5.0, -32.0)
    assert (cap, e, a) == (5.0, 0.0, 0.0)


def _mini(n_ports, sessions, imin=0.0, T=8):
    topo = Topology.uniform(C=1, n_ports=n_ports, Tr=1, T=T, imin=imin)
    S = len(sessions)
    z = lambda v: np.full(S, v, dtype=np.float64)
    sess = dict(loc=np.zeros(S, np.int32), t_arr=np.array([s[0] for s in sessions], np.int32),
                t_dep=np.array([s[1] for s in sessions], np.int32), ev_phases=np.full(S, 3, np.int32),
                lut=np.full(S, -1, np.int32), cap0=np.array([s[2] for s in sessions], np.float64), B=z(50),
                pmax_ac=z(11), pmin_ac=z(0), pmax_dis=z(-11), pmin_dis=z(0), bmin=z(5), bmin_em=z(25),
                desired=z(50), ts=z(1), mult=z(1), eta_c=z(1), eta_d=z(1))
    sc = Scenario(charge_price=np.full(T, -0.05), discharge_price=np.full(T, 0.05), setpoint=np.zeros(T),
                  tr_infl=np.zeros((1, T)), tr_solar=np.zeros((1, T)), tr_max_power=np.full((1, T), 100.0),
                  tr_min_power=np.full((1, T), -100.0), tr_load_fc=np.zeros((1, T)), tr_pv_fc=np.zeros((1, T)),
                  dr_start=np.zeros((1, 1), np.int32), dr_end=np.zeros((1, 1), np.int32), dr_cap=np.zeros((1, 1)),
                  dr_count=np.zeros(1, np.int32), sessions=sess).normalise()
    return topo, sc


def test_known_answers_charger():
    """KA3 (two-port normalisation), KA4 (min-current gate), KA5 (empty-port zeroing), KA6 (departure order)."""
    topo, sc = _mini(2, [(1, 7, 20.0), (1, 7, 30.0)])
    env = OracleEnv(topo, sc, reward="profit_maximization")
    env.reset()
    env.step(np.zeros(2))                                    # EVs arrive at the end of step 0
    r = env.step(np.array([1.0, 1.0]))                       # KA3: normalised to .5/.5 -> 16 A each
    assert list(r["port_cap"]) == [22.75, 32.75]
    assert r["cs_power"][0] == 22.000000000000007 and r["cs_current"][0] == 31.75426480542943
    r = env.step(np.array([0.3, 0.3]))
    assert list(r["port_cap"]) == [24.42, 34.42] and r["cs_power"][0] == 13.302150202128969

    topo, sc = _mini(1, [(1, 7, 20.0)], imin=6.0)            # KA4
    env = OracleEnv(topo, sc)
    env.reset(); env.step(np.zeros(1))
    r = env.step(np.array([0.1]))
    assert r["port_cap"][0] == 20.0 and r["cs_power"][0] == 0.0

    topo, sc = _mini(2, [(1, 7, 20.0)])                      # KA5
    env = OracleEnv(topo, sc)
    env.reset(); env.step(np.zeros(2))
    r = env.step(np.array([0.5, 0.7]))
    assert list(r["actions_eff"]) == [0.5, 0.0] and r["invalid_actions"] == 1

    topo, sc = _mini(1, [(1, 3, 20.0)])                      # KA6: charged on steps 1,2,3 then leaves in step 3
    env = OracleEnv(topo, sc, reward="profit_maximization")
    env.reset(); env.step(np.zeros(1))
    caps = []
    for _ in range(3):
        r = env.step(np.ones(1))
        caps.append(r["port_cap"][0])
    assert caps == [22.75, 25.5, 28.26] and r["n_departed"] == 1 and r["port_session"][0] == -1
    assert r["dep_sat"][0] == 28.26 / 50


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_statistics_match_reference(name):
    """get_statistics (utils.py:12-123) incl. battery degradation (ev.py:442-521): 1e-10 relative
    (numpy sums pairwise, the oracle sequentially); AFAP bounds (ev.py:407-440) bit-exact."""
    from oracle.oracle import STAT_NAMES
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    env = OracleEnv(pack.topo, pack.scenarios[0], reward=str(tr["reward_fn"]), state=str(tr["state_fn"]))
    env.reset()
    for t in range(tr["reward"].shape[0]):
        env.step(tr["actions"][t])
    st = env.statistics()
    for k in STAT_NAMES:
        ref = float(tr["stat_" + k])
        assert (np.isnan(ref) and np.isnan(st[k])) or st[k] == pytest.approx(ref, rel=1e-10, abs=1e-12), k
    assert np.array_equal(env.arr["ev_afap"][env.arr["ev_spawned"] > 0], tr["afap"])


@pytest.mark.parametrize("name", [n for n in golden_cases() if "roundrobin" in n or "calap" in n])
def test_oracle_agents_match_reference_agents(name):
    """oracle/agents.py emits, step by step, exactly the action vector the reference's RoundRobin /
    ChargeAsLateAsPossible (heuristics.py:7-150) produced on the same episode (bit-exact float64)."""
    from oracle.agents import OracleChargeAsLateAsPossible, OracleRoundRobin
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    env = OracleEnv(pack.topo, pack.scenarios[0], reward=str(tr["reward_fn"]), state=str(tr["state_fn"]))
    env.reset()
    agent = OracleRoundRobin(env) if "roundrobin" in name else OracleChargeAsLateAsPossible()
    nonzero = 0
    for t in range(tr["reward"].shape[0]):
        a = agent.get_action(env)
        assert np.array_equal(a, tr["actions"][t]), t
        nonzero += int(np.count_nonzero(a))
        env.step(a)
    assert nonzero > 0
