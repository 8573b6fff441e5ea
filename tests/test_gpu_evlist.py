"""GPU parity of the event-driven step kernel (ev2gym_b200/csrc/ev2b_evlist.cuh, handles created under
EV2B_KERNEL=evlist) through the C ABI: against the C oracle on seeded synthetic scenarios, against the recorded
reference traces, and against step_kernel at a BASELINE-sized batch.

Bars: battery level / occupancy / done / counts: BIT-EXACT; float64 outputs 1e-9 relative (sums are formed in a
different fixed order than the reference's sequential loops); float32 outputs (obs, cs_power, cs_current) 1e-5.
"""
import numpy as np
import pytest

from conftest import GOLDEN, golden_cases

pytestmark = pytest.mark.gpu

OUT = ("reward", "status", "obs", "tr_power", "tr_overload", "cs_power", "cs_current", "total_costs", "action_mask")


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def _engine(topo, bank, E, reward, state, outputs=OUT):
    from ev2gym_b200.engine import BatchedEngine
    eng = BatchedEngine(topo, E, reward=reward, state=state, outputs=outputs)
    eng.load_scenarios(bank)
    return eng


SHAPES = [  # C, n_ports, Tr, E, reward, state, action dtype
    (25, 1, 1, 70, "SquaredTrackingErrorReward", "PublicPST", "float32"),
    (100, 2, 5, 33, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float32"),
    (250, 1, 1, 9, "profit_maximization", "V2G_profit_max", "float64"),
    (7, 3, 2, 50, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", "float64"),
    (40, 2, 5, 21, "V2G_profitmaxV2", "V2G_profit_max_loads", "float32"),
    (25, 1, 1, 40, "SqTrError_TrPenalty_UserIncentives", "PublicPST", "float64"),
    (30, 2, 3, 17, "SquaredTrackingErrorRewardWithPenalty", "PublicPST", "float32"),
    (30, 1, 2, 19, "pst_V2G_profitmaxV2", "V2G_profit_max", "float32"),
]


@pytest.mark.parametrize("G", [1, 2, 4])
@pytest.mark.parametrize("C,n,Tr,E,reward,state,adt", SHAPES)
def test_evlist_matches_oracle_on_synthetic(C, n, Tr, E, reward, state, adt, G, monkeypatch):
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    monkeypatch.setenv("EV2B_EVL_G", str(G))
    topo = Topology.uniform(C=C, n_ports=n, Tr=Tr, T=64, imin=6.0 if n == 3 else 0.0)
    bank = sample_bank(topo, 5, seed=C + n, min_stay=5)
    scn_ids = [(3 * e + 1) % 5 for e in range(E)]
    eng = _engine(topo, bank, E, reward, state)
    obs0 = eng.reset(scn_ids=scn_ids).cpu().numpy()
    orc = OracleBatch(topo, [bank[i] for i in scn_ids], reward=reward, state=state)
    assert _close(obs0, orc.reset(), 1e-5, 1e-5)
    st = eng.state_tensors()
    rng = np.random.default_rng(99)
    for t in range(topo.T):
        a = rng.uniform(-1, 1, (E, topo.P))
        a[rng.random((E, topo.P)) < 0.1] = 0.0
        if t % 9 == 4:
            a[:] = 1.0
        a = a.astype(adt)
        out = {k: v.cpu().numpy() for k, v in eng.step(torch.tensor(a, device="cuda")).items()}
        orc.step(a.astype(np.float64))
        occ = orc.arr["port_session"] >= 0
        cap = st["port_cap"].cpu().numpy()
        assert np.array_equal(cap[occ], orc.arr["port_cap"][occ]), (t, "cap")            # bit exact
        hot = eng.decode_hot(st["port_hot"].cpu().numpy())
        live = orc.done == 0
        occ_dev = (hot["t_arr"] <= t + 1) & (t + 1 <= hot["t_dep"])
        assert np.array_equal(occ_dev[live], occ[live]), (t, "occupancy")
        assert np.array_equal(out["action_mask"] > 0, occ), (t, "action mask")
        assert _close(out["reward"], orc.reward, 1e-9, 1e-9), t
        assert _close(out["total_costs"], [o.total_costs for o in orc.outs], 1e-9, 1e-12), t
        assert _close(out["tr_power"], orc.o["tr_power"][:, :Tr], 1e-9, 1e-9), t
        assert _close(out["tr_overload"], orc.o["tr_overload"][:, :Tr], 1e-9, 1e-9), t
        assert _close(out["cs_power"], orc.o["cs_power"], 1e-5, 1e-6), t
        assert _close(out["cs_current"], orc.o["cs_current"], 1e-5, 1e-6), t
        assert _close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), t
        assert np.array_equal((out["status"] & 1) > 0, orc.done > 0), t
        ovf = np.array([o.error == 1 for o in orc.outs])
        assert np.array_equal((out["status"] & 2) > 0, ovf), (t, "amps overflow flag")
    k = eng.kpis()
    assert _close(k["total_reward"], [s.total_reward for s in orc.states], 1e-9, 1e-9)
    assert np.array_equal(k["total_evs_spawned"], [float(s.total_evs_spawned) for s in orc.states])
    assert eng.kernel_launches() == (0, topo.T, 0)      # every launch took the event-driven kernel
    eng.close()


def _lean_golden():
    out = []
    for name in golden_cases():
        tr = np.load(f"{GOLDEN}/{name}.trace.npz")
        if not name.startswith("grid"):          # every stock reward / state function that needs no distribution grid
            out.append(name)
    return out


@pytest.mark.parametrize("name", _lean_golden())
def test_evlist_matches_reference_trace(name, monkeypatch):
    import torch
    from ev2gym_b200.scenario import ScenarioPack
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    topo = pack.topo
    E = 3
    eng = _engine(topo, pack.scenarios, E, str(tr["reward_fn"]), str(tr["state_fn"]))
    obs0 = eng.reset().cpu().numpy()
    assert _close(obs0[0], tr["obs0"], 1e-5, 1e-6)
    st = eng.state_tensors()
    T = tr["reward"].shape[0]
    for t in range(T):
        a = torch.tensor(np.tile(tr["actions"][t], (E, 1)), dtype=torch.float64, device="cuda")
        out = {k: v.cpu().numpy() for k, v in eng.step(a).items()}
        for e in (0, E - 1):
            occ = tr["action_mask"][t] > 0
            cap = st["port_cap"][e].cpu().numpy()
            assert np.array_equal(cap[occ], tr["cap"][t][occ]), (t, "cap")
            assert np.array_equal(out["action_mask"][e] > 0, occ), (t, "action mask")
            hot = eng.decode_hot(st["port_hot"][e].cpu().numpy())
            assert np.array_equal(hot["t_arr"][occ], tr["port_t_arr"][t][occ]), (t, "arrival index")
            assert _close(out["reward"][e], tr["reward"][t], 1e-9, 1e-9), (t, out["reward"][e], tr["reward"][t])
            assert _close(out["total_costs"][e], tr["total_costs"][t], 1e-9, 1e-12)
            assert _close(out["tr_power"][e], tr["tr_power"][t], 1e-9, 1e-9)
            assert _close(out["tr_overload"][e], tr["tr_overload"][t], 1e-9, 1e-9)
            assert _close(out["cs_power"][e], tr["cs_power"][t], 1e-5, 1e-6)
            assert _close(out["cs_current"][e], tr["cs_current"][t], 1e-5, 1e-6)
            assert _close(out["obs"][e], tr["obs"][t], 1e-5, 1e-5), (t, "obs")
            assert bool(out["status"][e] & 1) == bool(tr["done"][t])
    k = eng.kpis()
    assert k["total_reward"][0] == pytest.approx(float(tr["total_reward"]), rel=1e-9, abs=1e-9)
    assert k["total_ev_served"][0] == float(tr["stat_total_ev_served"])
    assert eng.kernel_launches() == (0, T, 0)
    eng.close()


@pytest.mark.parametrize("G", [1, 4])
def test_evlist_equals_step_kernel_at_c3_size(G, monkeypatch):
    """A whole c3-shaped episode (512 envs x 100 chargers x 2 ports, 5 transformers) on both kernels: identical
    hot / cap / exch arrays and observations, rewards and KPI sums within 1e-9; then mixed use on one handle
    (per-port outputs force step_kernel for some launches, the list is re-derived afterwards)."""
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=100, n_ports=2, Tr=5, T=112)
    bank = sample_bank(topo, 16, seed=21, min_stay=5)
    E = 512
    gen = torch.Generator(device="cuda")
    res = {}
    for kn in ("percharger", "evlist", "mixed"):
        monkeypatch.setenv("EV2B_KERNEL", "percharger" if kn == "percharger" else "evlist")
        monkeypatch.setenv("EV2B_EVL_G", str(G))
        monkeypatch.setenv("EV2B_EVL_MIX", "1" if kn == "mixed" else "0")   # test-only: port_energy launches take step_kernel
        eng = _engine(topo, bank, E, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads",
                      outputs=("reward", "status", "obs"))
        eng.reset()
        gen.manual_seed(5)
        rew = []
        for t in range(topo.T):
            a = torch.rand((E, topo.P), device="cuda", generator=gen) * 2.0 - 1.0
            if kn == "mixed" and t % 7 in (2, 3):
                eng.set_outputs(("reward", "status", "obs", "port_energy"))
                eng.step(a)
                eng.set_outputs(("reward", "status", "obs"))
                rew.append(eng.out["reward"].clone() * 0)      # (outputs were re-allocated: rewards of these steps not compared)
                continue
            rew.append(eng.step(a)["reward"].clone())
        st = eng.state_tensors()
        res[kn] = (st["port_hot"].cpu().numpy().copy(), st["port_cap"].cpu().numpy().copy(),
                   st["port_exch"].cpu().numpy().copy(), torch.stack(rew).cpu().numpy(), st["env_kpi"].cpu().numpy().copy(),
                   eng.kernel_launches())
        assert bool((eng.out["status"] & 1).all())
        eng.close()
    a, b, m = res["percharger"], res["evlist"], res["mixed"]
    assert a[5] == (topo.T, 0, 0) and b[5] == (0, topo.T, 0)
    assert m[5][0] > 0 and m[5][1] > 0 and m[5][2] > 0
    for i in range(3):
        assert np.array_equal(a[i], b[i]), i
        assert np.array_equal(a[i], m[i]), i
    assert _close(b[3], a[3], 1e-9, 1e-9)
    assert _close(b[4], a[4], 1e-9, 1e-9) and _close(m[4], a[4], 1e-9, 1e-9)


@pytest.mark.parametrize("adt", ["float32", "float64"])
def test_unaligned_action_and_observation_buffers(adt, monkeypatch):
    """The event-driven kernel reads both actions of a charger with one load and writes an EV's observation tuple / the
    series values with 8-byte stores when the caller's buffers allow it.  Buffers that start one element into an allocation
    take the scalar paths and must give the same episode."""
    import torch
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    topo = Topology.uniform(C=40, n_ports=2, Tr=5, T=40)
    bank = sample_bank(topo, 4, seed=3, min_stay=5)
    E = 37
    a_eng = _engine(topo, bank, E, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", outputs=("reward", "status", "obs"))
    b_eng = _engine(topo, bank, E, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", outputs=("reward", "status", "obs"))
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if adt == "float32" else torch.float64
    raw_act = torch.zeros(E * topo.P + 1, dtype=tdt, device=dev)
    act_odd = raw_act[1:].view(E, topo.P)                        # one element into the allocation: not on a pair boundary
    raw_obs = torch.zeros(E * a_eng.D + 1, dtype=torch.float32, device=dev)
    obs_odd = raw_obs[1:].view(E, a_eng.D)                       # 4 bytes off an 8-byte boundary
    assert act_odd.data_ptr() % (2 * act_odd.element_size()) != 0 and obs_odd.data_ptr() % 8 != 0
    b_eng.out["obs"] = obs_odd                                   # (the engine's own buffer is replaced before the first reset)
    b_eng._so.obs = obs_odd.data_ptr()
    a_eng.reset(); b_eng.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    for t in range(topo.T):
        a = (torch.rand((E, topo.P), device=dev, generator=gen) * 2 - 1).to(tdt)
        act_odd.copy_(a)
        oa = a_eng.step(a)
        ob = b_eng.step(act_odd)
        assert torch.equal(oa["reward"], ob["reward"]) and torch.equal(oa["status"], ob["status"]), t
        assert torch.equal(oa["obs"], obs_odd), t
    assert torch.equal(a_eng.state_tensors()["port_cap"], b_eng.state_tensors()["port_cap"])
    a_eng.close(); b_eng.close()
