"""GPU parity: the CUDA step engine (through the C ABI, libev2b.so) against
  (1) traces recorded from the Python reference (tests/golden), and
  (2) the C oracle on seeded synthetic scenarios at sizes the oracle finishes in seconds.

Bars (written where they are asserted):
  - battery level / arrival-departure indexing / action mask / done flags: BIT-EXACT
  - float64 outputs (reward, total_costs, tr_power, tr_overload): 1e-9 relative
    (the device sums chargers in a fixed tree order, the reference sequentially)
  - float32 outputs (obs, cs_power, cs_current): 1e-5 relative (north_star tolerance; fp32 storage)
"""
import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

pytestmark = pytest.mark.gpu


def _engine(topo, scenarios, n_envs, reward, state, outputs, stats=False):
    from ev2gym_b200.engine import BatchedEngine
    eng = BatchedEngine(topo, n_envs, reward=reward, state=state, outputs=outputs, stats=stats)
    eng.load_scenarios(scenarios)
    return eng


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


ALL_OUT = ("reward", "status", "obs", "cs_power", "cs_current", "tr_power", "tr_overload", "total_costs",
           "action_mask", "dep_sat", "port_energy")


@pytest.mark.parametrize("epb", [0, 3])
@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_reference_trace(name, epb, monkeypatch):
    import torch
    if epb:                               # small batches get one env per CTA by default: force the multi-env CTA layout too
        monkeypatch.setenv("EV2B_EPB", str(epb))
    from ev2gym_b200.scenario import ScenarioPack
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    topo = pack.topo
    E = 3                                 # 3 replicas of the same episode: exercises multi-env CTAs
    grid = topo.n_bus > 0
    eng = _engine(topo, pack.scenarios, E, str(tr["reward_fn"]), str(tr["state_fn"]),
                  ALL_OUT + (("node_voltage",) if grid else ()), stats=True)
    obs0 = eng.reset().cpu().numpy()
    assert _close(obs0[0], tr["obs0"], 1e-5, 1e-6) and np.array_equal(obs0[0], obs0[2])
    st = eng.state_tensors()
    T = tr["reward"].shape[0]
    for t in range(T):
        a = torch.tensor(np.tile(tr["actions"][t], (E, 1)), dtype=torch.float64, device="cuda")
        out = {k: v.cpu().numpy() for k, v in eng.step(a).items()}
        for e in (0, E - 1):
            occ = out["action_mask"][e] > 0
            assert np.array_equal(occ.astype(np.float64), tr["action_mask"][t]), (t, "mask")
            cap = st["port_cap"][e].cpu().numpy()
            # battery level: bit exact (fp64, same operation order, no FMA)
            assert np.array_equal(cap[occ], tr["cap"][t][occ]), (t, "cap", cap[occ], tr["cap"][t][occ])
            hot = eng.decode_hot(st["port_hot"][e].cpu().numpy())
            assert np.array_equal(hot["t_arr"][occ], tr["port_t_arr"][t][occ]), (t, "arrival index")
            assert _close(out["reward"][e], tr["reward"][t], 1e-9, 1e-9), (t, out["reward"][e], tr["reward"][t])
            assert _close(out["total_costs"][e], tr["total_costs"][t], 1e-9, 1e-12)
            assert _close(out["tr_power"][e], tr["tr_power"][t], 1e-9, 1e-9)
            assert _close(out["tr_overload"][e], tr["tr_overload"][t], 1e-9, 1e-9)
            assert _close(out["cs_power"][e], tr["cs_power"][t], 1e-5, 1e-6)
            assert _close(out["cs_current"][e], tr["cs_current"][t], 1e-5, 1e-6)
            assert _close(out["obs"][e], tr["obs"][t], 1e-5, 1e-5), (t, "obs",
                                                                     np.abs(out["obs"][e] - tr["obs"][t]).max())
            if grid:      # Laurent power flow: |V| per node, slack first (ev2gym_env.py:397)
                assert _close(out["node_voltage"][e], tr["node_voltage"][:, t], 1e-9, 1e-12), (t, "node voltage")
            assert bool(out["status"][e] & 1) == bool(tr["done"][t])
            assert np.nansum(out["dep_sat"][e]) == pytest.approx(tr["sat_sum"][t], rel=1e-6, abs=1e-6)
            assert np.count_nonzero(~np.isnan(out["dep_sat"][e])) == tr["n_departed"][t]
    k = eng.kpis()
    assert k["total_reward"][0] == pytest.approx(float(tr["total_reward"]), rel=1e-9, abs=1e-9)
    assert k["total_ev_served"][0] == float(tr["stat_total_ev_served"])
    assert k["total_profits"][0] == pytest.approx(float(tr["stat_total_profits"]), rel=1e-9, abs=1e-9)
    assert k["total_energy_charged"][0] == pytest.approx(float(tr["stat_total_energy_charged"]), rel=1e-9)
    assert k["total_transformer_overload"][0] == pytest.approx(float(tr["stat_total_transformer_overload"]),
                                                               rel=1e-9, abs=1e-9)
    assert k["tracking_error"][0] == pytest.approx(float(tr["stat_tracking_error"]), rel=1e-9, abs=1e-9)
    # get_statistics(env) incl. battery degradation and AFAP-normalised energy satisfaction: 1e-9 relative
    from ev2gym_b200.engine import STAT_NAMES
    st = eng.episode_stats()
    for n in STAT_NAMES:
        ref = float(tr["stat_" + n])
        for e in (0, E - 1):
            assert (np.isnan(ref) and np.isnan(st[n][e])) or st[n][e] == pytest.approx(ref, rel=1e-9, abs=1e-12), n
    # stepping a finished env is a no-op flagged WAS_DONE (reference: AssertionError, ev2gym_env.py:343)
    out = eng.step(a)
    assert int(out["status"][0].item()) & 4 and float(out["reward"][0].item()) == 0.0


SHAPES = [  # C, n_ports, Tr, E, reward, state, action dtype
    (25, 1, 1, 37, "SquaredTrackingErrorReward", "PublicPST", "float32"),
    (100, 2, 5, 9, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float32"),
    (250, 1, 1, 4, "profit_maximization", "V2G_profit_max", "float64"),
    (7, 3, 2, 50, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float64"),
    (300, 1, 20, 3, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float32"),
]


@pytest.mark.parametrize("C,n,Tr,E,reward,state,adt", SHAPES)
def test_cuda_matches_oracle_on_synthetic(C, n, Tr, E, reward, state, adt, monkeypatch):
    if C * 4 <= 1024:
        monkeypatch.setenv("EV2B_EPB", "4")   # several envs per CTA (the default for big batches of small envs)
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch
    topo = Topology.uniform(C=C, n_ports=n, Tr=Tr, T=64, imin=6.0 if n == 3 else 0.0)
    bank = sample_bank(topo, 5, seed=C + n, min_stay=5)
    scn_ids = [(3 * e + 1) % 5 for e in range(E)]
    eng = _engine(topo, bank, E, reward, state, ALL_OUT, stats=True)
    obs0 = eng.reset(scn_ids=scn_ids).cpu().numpy()
    orc = OracleBatch(topo, [bank[i] for i in scn_ids], reward=reward, state=state)
    assert _close(obs0, orc.reset(), 1e-5, 1e-5)
    st = eng.state_tensors()
    rng = np.random.default_rng(99)
    for t in range(topo.T):
        a = rng.uniform(-1, 1, (E, topo.P))
        a[rng.random((E, topo.P)) < 0.1] = 0.0
        a = a.astype(adt)
        out = {k: v.cpu().numpy() for k, v in eng.step(torch.tensor(a, device="cuda")).items()}
        orc.step(a.astype(np.float64))
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(out["action_mask"] > 0, occ), t
        cap = st["port_cap"].cpu().numpy()
        assert np.array_equal(cap[occ], orc.arr["port_cap"][occ]), (t, "cap")            # bit exact
        assert _close(out["reward"], orc.reward, 1e-9, 1e-9), t
        assert _close(out["tr_power"], orc.o["tr_power"][:, :Tr], 1e-9, 1e-9), t
        assert _close(out["tr_overload"], orc.o["tr_overload"][:, :Tr], 1e-9, 1e-9), t
        assert _close(out["cs_power"], orc.o["cs_power"], 1e-5, 1e-6), t
        assert _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), t
        assert np.array_equal((out["status"] & 1) > 0, orc.done > 0), t
        ovf = np.array([o.error == 1 for o in orc.outs])
        assert np.array_equal((out["status"] & 2) > 0, ovf), (t, "amps overflow flag")
    assert _close(eng.kpis()["total_reward"], [s.total_reward for s in orc.states], 1e-9, 1e-9)
    st, ost = eng.episode_stats(), orc.statistics()
    for n in st:
        both_nan = np.isnan(st[n]) & np.isnan(ost[n])
        assert _close(np.where(both_nan, 0, st[n]), np.where(both_nan, 0, ost[n]), 1e-9, 1e-12), n


def test_ragged_ports_and_reset_done():
    """Chargers with different port counts / currents (topology-file style) + device-side auto reset."""
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch
    topo = Topology(cs_n_ports=[2, 1, 3, 2, 1, 4], cs_tr=[0, 0, 1, 1, 2, 2], cs_imax=[56, 32, 56, 16, 32, 56],
                    cs_imin=[8, 0, 8, 0, 6, 8], cs_imax_dis=[0, -32, -56, 0, -32, -56], cs_imin_dis=[0] * 6,
                    cs_voltage=[230, 400, 230, 400, 400, 230], cs_phases=[3, 3, 1, 3, 3, 2], n_transformers=3,
                    tr_voltage=400 * 3 ** 0.5, timescale=15, sim_length=40, dr_steps_ahead=4)
    bank = sample_bank(topo, 6, seed=5, min_stay=4)
    E = 11
    eng = _engine(topo, bank, E, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", ALL_OUT)
    eng.reset()
    ids = [e % 6 for e in range(E)]
    orc = OracleBatch(topo, [bank[i] for i in ids], reward="ProfitMax_TrPenalty_UserIncentives",
                      state="V2G_profit_max_loads")
    orc.reset()
    rng = np.random.default_rng(3)
    st = eng.state_tensors()
    for ep in range(2):
        for t in range(topo.T):
            a = rng.uniform(-0.4, 1, (E, topo.P))
            out = {k: v.cpu().numpy() for k, v in eng.step(torch.tensor(a, device="cuda")).items()}
            orc.step(a)
            occ = orc.arr["port_session"] >= 0
            assert np.array_equal(st["port_cap"].cpu().numpy()[occ], orc.arr["port_cap"][occ])
            assert _close(out["reward"], orc.reward, 1e-9, 1e-9)
            assert _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5)
        assert np.all(out["status"] & 1)
        obs = eng.reset_done().cpu().numpy()           # scenario id advances by E modulo the bank size
        ids = [(i + E) % 6 for i in ids]
        assert np.array_equal(st["env_scn"].cpu().numpy(), ids)
        orc = OracleBatch(topo, [bank[i] for i in ids], reward="ProfitMax_TrPenalty_UserIncentives",
                          state="V2G_profit_max_loads")
        assert _close(obs, orc.reset(), 1e-5, 1e-5)


@pytest.mark.parametrize("pinned", [False, True])
def test_step_host_matches_device_path(pinned):
    """ev2b_step_host == ev2b_step, with pageable host arrays (plain stream calls) and with pinned ones (the call is
    same path, the copies then run at full PCIe rate; two action buffers alternate)."""
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=20, n_ports=2, Tr=2, T=30)
    bank = sample_bank(topo, 3, seed=8, min_stay=4)
    E = 16
    e1 = _engine(topo, bank, E, "profit_maximization", "V2G_profit_max", ("reward", "status", "obs"))
    e2 = _engine(topo, bank, E, "profit_maximization", "V2G_profit_max", ("reward", "status", "obs"))
    e1.reset(); e2.reset()
    rng = np.random.default_rng(1)
    if pinned:
        keep = [torch.zeros(E, dtype=torch.float64).pin_memory(), torch.zeros(E, dtype=torch.int32).pin_memory(),
                torch.zeros((E, e1.D), dtype=torch.float32).pin_memory(),
                torch.zeros((E, topo.P), dtype=torch.float32).pin_memory(), torch.zeros((E, topo.P), dtype=torch.float32).pin_memory()]
        rew, stt, obs = keep[0].numpy(), keep[1].numpy().view(np.uint32), keep[2].numpy()
        abuf = [keep[3].numpy(), keep[4].numpy()]
    else:
        rew, stt, obs = np.zeros(E), np.zeros(E, dtype=np.uint32), np.zeros((E, e1.D), dtype=np.float32)
    for t in range(topo.T):
        a = rng.uniform(-1, 1, (E, topo.P)).astype(np.float32)
        if pinned:
            abuf[t % 2][...] = a
            a = abuf[t % 2]
        if t == 17:                      # a reset in the middle must not be hidden by the replayed graphs
            e1.reset(); e2.reset()
        o1 = e1.step(torch.tensor(a, device="cuda"))
        e2.step_host(a, rew, stt, obs)
        assert np.array_equal(o1["reward"].cpu().numpy(), rew)
        assert np.array_equal(o1["obs"].cpu().numpy(), obs)
        assert np.array_equal(o1["status"].cpu().numpy().astype(np.uint32), stt)


@pytest.mark.parametrize("agent", ["afap", "zero", "uniform", "external"])
def test_step_k_device_agents(agent):
    """ev2b_step_k: k steps per call with an on-device agent == stepping the oracle with the same actions."""
    import torch
    from ev2gym_b200.engine import BatchedEngine
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch
    topo = Topology.uniform(C=30, n_ports=2, Tr=3, T=48)
    bank = sample_bank(topo, 4, seed=21, min_stay=5)
    E, seed = 13, 0x1234ABCD5678
    rw, stf = "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads"
    eng = _engine(topo, bank, E, rw, stf, ("reward", "status", "obs"))
    eng.reset()
    orc = OracleBatch(topo, [bank[e % 4] for e in range(E)], reward=rw, state=stf)
    orc.reset()
    caps = eng.state_tensors()["port_cap"]
    rng = np.random.default_rng(2)
    t = 0
    for k in (1, 5, 16, 26):
        ext = rng.uniform(-1, 1, (k, E, topo.P))
        for i in range(k):
            a = {"afap": np.ones((E, topo.P)), "zero": np.zeros((E, topo.P)), "external": ext[i],
                 "uniform": BatchedEngine.uniform_agent_actions(seed, E, topo.P, t + i, -1.0)}[agent]
            orc.step(a)
        out = eng.step_k(k, agent, actions_k=torch.tensor(ext, device="cuda") if agent == "external" else None,
                         seed=seed)
        t += k
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(caps.cpu().numpy()[occ], orc.arr["port_cap"][occ]), (agent, t)
        assert _close(out["reward"].cpu().numpy(), orc.reward, 1e-9, 1e-9)
        assert _close(out["obs"].cpu().numpy(), orc.o["obs"][:, :eng.D], 1e-5, 1e-5)
    assert t == topo.T and bool((out["status"] & 1).all())
    assert _close(eng.kpis()["total_reward"], [s.total_reward for s in orc.states], 1e-9, 1e-9)


def test_grid_power_flow_matches_oracle_on_synthetic(monkeypatch):
    monkeypatch.setenv("EV2B_EPB", "3")
    """Synthetic 20-bus feeder, 2 ports per charger, several envs per CTA: voltages / grid rewards / grid state."""
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import add_grid, sample_bank
    from oracle.oracle import OracleBatch
    topo = Topology.uniform(C=50, n_ports=2, Tr=20, T=40)
    bank = sample_bank(topo, 3, seed=31, min_stay=4, loads=False)
    add_grid(topo, bank, seed=1)
    E = 7
    for rw in ("V2G_grid_full_reward", "V2G_grid_simple_reward"):
        eng = _engine(topo, bank, E, rw, "V2G_grid_state", ("reward", "status", "obs", "node_voltage"))
        obs0 = eng.reset().cpu().numpy()
        orc = OracleBatch(topo, [bank[e % 3] for e in range(E)], reward=rw, state="V2G_grid_state")
        assert _close(obs0, orc.reset(), 1e-5, 1e-5)
        rng = np.random.default_rng(4)
        for t in range(topo.T):
            a = rng.uniform(-1, 1, (E, topo.P))
            out = {k: v.cpu().numpy() for k, v in eng.step(torch.tensor(a, device="cuda")).items()}
            orc.step(a)
            assert _close(out["node_voltage"], orc.o["node_vm"], 1e-9, 1e-12), t
            assert _close(out["reward"], orc.reward, 1e-9, 1e-9), t
            assert _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), t
        eng.close()


@pytest.mark.parametrize("name", [n for n in golden_cases() if "roundrobin" in n or "calap" in n])
def test_device_agents_match_reference_agents(name):
    """ev2b_agent_actions == the action vector the reference's RoundRobin / ChargeAsLateAsPossible emitted at every step
    of the recorded episode (float64, bit-exact: queue order, fractional last EV, CALAP's ceil)."""
    from ev2gym_b200.scenario import ScenarioPack
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    E = 3
    eng = _engine(pack.topo, pack.scenarios, E, str(tr["reward_fn"]), str(tr["state_fn"]), ("reward", "status"))
    eng.reset()
    kind = "roundrobin" if "roundrobin" in name else "calap"
    for t in range(tr["reward"].shape[0]):
        a = eng.agent_actions(kind)
        got = a.cpu().numpy()
        assert np.array_equal(got[0], tr["actions"][t]) and np.array_equal(got[E - 1], tr["actions"][t]), t
        out = eng.step(a)
        assert _close(out["reward"][0].item(), tr["reward"][t], 1e-9, 1e-9), t
    assert bool((out["status"] & 1).all())
    assert not eng.agent_actions(kind).any()            # finished envs: zeros


@pytest.mark.parametrize("agent,C,n", [("roundrobin", 30, 2), ("calap", 30, 2), ("roundrobin", 17, 3), ("roundrobin", 600, 2)])
def test_step_k_tensor_agents(agent, C, n):
    """ev2b_step_k with the queue-based / whole-env agents == oracle env driven by oracle/agents.py, including the
    queue reset at ev2b_reset and ports beyond one CTA pass (P = 1200)."""
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.agents import OracleChargeAsLateAsPossible, OracleRoundRobin
    from oracle.oracle import OracleBatch
    T = 12 if C > 100 else 40
    topo = Topology.uniform(C=C, n_ports=n, Tr=2, T=T, v2g_enabled=False)
    bank = sample_bank(topo, 3, seed=33, min_stay=4, occupancy=0.8)
    for sc in bank:                       # setpoints that serve a fraction of the waiting EVs (fractional last share)
        sc.setpoint = np.round(sc.setpoint * (0.02 if C > 100 else 0.35), 3)
        sc.normalise()
    E = 4 if C > 100 else 7
    rw, stf = "SquaredTrackingErrorReward", "PublicPST"
    eng = _engine(topo, bank, E, rw, stf, ("reward", "status", "obs"))
    orc = OracleBatch(topo, [bank[e % 3] for e in range(E)], reward=rw, state=stf)
    caps = eng.state_tensors()["port_cap"]
    for episode in range(2):
        eng.reset(); orc.reset()
        views = [orc.env_view(e) for e in range(E)]
        agents = [OracleRoundRobin(v) if agent == "roundrobin" else OracleChargeAsLateAsPossible() for v in views]
        t, served = 0, 0
        for k in ((3, 9) if C > 100 else (1, 7, 20, 12)):
            for i in range(k):
                a = np.stack([ag.get_action(v) for ag, v in zip(agents, views)])
                served += int(np.count_nonzero(a))
                orc.step(a)
            out = eng.step_k(k, agent)
            t += k
            occ = orc.arr["port_session"] >= 0
            assert np.array_equal(caps.cpu().numpy()[occ], orc.arr["port_cap"][occ]), (agent, episode, t)
            assert _close(out["reward"].cpu().numpy(), orc.reward, 1e-9, 1e-9)
            assert _close(out["obs"].cpu().numpy(), orc.o["obs"][:, :eng.D], 1e-5, 1e-5)
        assert t == topo.T and served > 0
