"""Device-side scenario sampling (ev2b_set_spawn_tables / ev2b_resample_sessions; ev2gym_b200/csrc/ev2b_spawn.cuh):
the GPU port of the reference's EV_spawner + spawn_single_EV (ev2gym/utilities/utils.py:477-557, 177-345).

The reference draws from numpy's global Mersenne-Twister stream, so a draw-for-draw comparison is impossible; parity is
  (1) EXACT for the structural rules: arrival window (first arrival at step 3, none after T - min_stay - 1), port rest
      of two steps (utils.py:534-536), t_dep = int(stay + t + 3) >= t_arr + min_stay + 1, nothing after the episode end
      (:254-256), workplace opening hours (:509-520), first-free-port placement (ev_charger.py:273), battery level
      rules (:220-229), np.round(., 3) lattices of transition_soc / efficiencies, the model table;
  (2) DISTRIBUTIONAL against the reference itself: the scenario banks under ev2gym_b200/data were sampled by the
      unmodified reference (tools/make_golden.py --packs) -- sessions per scenario, arrival step, length of stay, battery
      level at arrival and EV model of the device sample must match them (two-sample Kolmogorov-Smirnov statistic /
      total-variation distance below the bar noted at each assertion);
  (3) the engine steps an episode on the sampled sessions exactly like the oracle does on the same sessions read back.

CPU: the SIMT-emulated build (small sample).  GPU (`-m gpu`): the real library, 8x larger sample, tighter bars."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
DATA = os.path.join(ROOT, "ev2gym_b200", "data")

CASES = [("c3_v2gloads_c100n2tr5", "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads"),
         ("c2_publicpst_c25", "SquaredTrackingErrorReward", "PublicPST"),
         ("c4_v2gprofitmax_c250", "profit_maximization", "V2G_profit_max")]


def _ks(a, b):
    """Two-sample Kolmogorov-Smirnov statistic."""
    a, b = np.sort(np.asarray(a, dtype=np.float64)), np.sort(np.asarray(b, dtype=np.float64))
    grid = np.concatenate([a, b])
    return float(np.max(np.abs(np.searchsorted(a, grid, side="right") / len(a) - np.searchsorted(b, grid, side="right") / len(b))))


def _load(name):
    from ev2gym_b200.scenario import ScenarioPack, SpawnTables
    return ScenarioPack.load(os.path.join(DATA, name + ".npz")), SpawnTables.load(os.path.join(DATA, "spawn_" + name + ".npz"))


def _check_sample(make_engine, name, reward, state, n_rounds, ks_bar, steps_vs_oracle):
    from ev2gym_b200.scenario import Scenario
    from oracle.oracle import OracleBatch
    pack, tab = _load(name)
    topo, S = pack.topo, len(pack)
    T, P = topo.T, topo.P
    eng = make_engine(topo, S, reward, state)
    eng.set_spawn_tables(tab)
    eng.load_scenarios(pack.scenarios)
    ref = {k: np.concatenate([sc.sessions[k] for sc in pack.scenarios]) for k in ("t_arr", "t_dep", "cap0", "B", "pmax_ac")}
    ref_n = np.array([sc.n_sessions for sc in pack.scenarios], dtype=np.float64)
    dev = {k: [] for k in ("t_arr", "t_dep", "cap0", "B", "pmax_ac", "model")}
    dev_n, first = [], None
    for r in range(n_rounds):
        eng.resample_sessions(seed=1234 + r)
        for s in range(S):
            d = eng.read_sessions(s)
            if first is None:
                first = d
            n = len(d["port"])
            dev_n.append(n)
            wd, hh, mm = (int(x) for x in tab.start[s % len(tab.start)])
            # ---- (1) structural rules, exact
            assert np.all(np.diff(d["t_arr"]) >= 0), "arrival order"
            if n:
                assert d["t_arr"].min() >= 3 and d["t_arr"].max() <= T - tab.min_stay_steps - 1          # utils.py:504, 316
                assert np.all(d["t_dep"] - d["t_arr"] >= tab.min_stay_steps + 1)                        # :251-252, 316-318
                if tab.empty_ports_at_end:
                    assert np.all(d["t_dep"] <= T - 2)                                                   # :254-256
                B = tab.model_B[d["model"]]
                assert np.all((d["cap0"] >= 0.0) & (d["cap0"] <= B))                                     # :220-229
                assert np.all(d["cap0"] <= np.maximum(tab.desired_frac * B, B - 1))
                small = d["cap0"] < tab.min_battery_capacity                                             # :228-229
                assert not np.any(small & (B > 2 * tab.min_battery_capacity))
                if tab.heterogeneous:
                    assert np.all((d["ts"] >= 0.7 - 1e-9) & (d["ts"] <= 0.9)), "transition_soc in [0.7, 0.9]"   # :309-310
                    assert np.allclose(d["ts"] * 1000, np.rint(d["ts"] * 1000), atol=1e-9), "np.round(., 3) lattice"
                    scalar = tab.model_lut[d["model"]] < 0
                    assert np.all(np.isnan(d["eta_c"][~scalar])) and np.all((d["eta_c"][scalar] >= 0.95 - 1e-9) & (d["eta_c"][scalar] <= 1.0))
                for port in np.unique(d["port"]):                      # a port's sessions never overlap
                    m = d["port"] == port
                    assert np.all(d["t_arr"][m][1:] > d["t_dep"][m][:-1]), (s, port)
                if tab.workplace:                                       # closed before 6 h, after 18 h, at weekends  :509-520
                    mins = hh * 60 + mm + (d["t_arr"] - 1 - 2) * topo.timescale
                    hour, day = (mins // 60) % 24, (wd + mins // 1440) % 7
                    assert np.all((hour >= 6) & (hour <= 18) & (day < 5))
                # first-free-port placement: replaying the rule over (charger, arrival order) reproduces the ports
                from ev2gym_b200.scenario import assign_ports
                loc = np.searchsorted(topo.cs_port_off, d["port"], side="right") - 1
                assert np.array_equal(assign_ports(topo, d["t_arr"], d["t_dep"], loc), d["port"])
            for k in ("t_arr", "t_dep", "cap0", "model"):
                dev[k].append(d[k])
            dev["B"].append(tab.model_B[d["model"]])
            dev["pmax_ac"].append(tab.model_pmax_ac[d["model"]])
    dev = {k: np.concatenate(v) for k, v in dev.items()}
    dev_n = np.array(dev_n, dtype=np.float64)
    # ---- (2) distributions against the reference-sampled bank
    assert abs(dev_n.mean() - ref_n.mean()) <= 4.0 * ref_n.std() / np.sqrt(len(ref_n)) + 0.03 * ref_n.mean(), \
        ("sessions per scenario", dev_n.mean(), ref_n.mean())
    assert _ks(dev["t_arr"], ref["t_arr"]) < ks_bar, ("arrival step", _ks(dev["t_arr"], ref["t_arr"]))
    assert _ks(dev["t_dep"] - dev["t_arr"], ref["t_dep"] - ref["t_arr"]) < ks_bar, "length of stay"
    assert _ks(dev["cap0"] / dev["B"], ref["cap0"] / ref["B"]) < ks_bar, ("state of charge at arrival", _ks(dev["cap0"] / dev["B"], ref["cap0"] / ref["B"]))
    for col in ("B", "pmax_ac"):                                       # EV model mix: total-variation distance of the histograms
        vals = np.unique(np.concatenate([dev[col], ref[col]]))
        hd = np.array([(dev[col] == v).mean() for v in vals]); hr = np.array([(ref[col] == v).mean() for v in vals])
        assert 0.5 * np.abs(hd - hr).sum() < 2.0 * ks_bar, (col, hd, hr)
    # ---- (2b) power setpoints regenerated from the sampled sessions (generate_power_setpoints, utils.py:664-757)
    ref_sp = np.stack([sc.setpoint[:T] for sc in pack.scenarios])
    dev_sp = np.stack([eng.read_setpoints(s) for s in range(S)])
    if tab.power_setpoint_enabled:
        assert np.all(dev_sp >= 0.0) and np.all(np.isfinite(dev_sp))
        # energy under the curve per scenario, shape over the day (mean profile), and the pooled values
        assert abs(dev_sp.sum(1).mean() - ref_sp.sum(1).mean()) <= 4.0 * ref_sp.sum(1).std() / np.sqrt(S) + 0.04 * ref_sp.sum(1).mean(), \
            ("setpoint energy", dev_sp.sum(1).mean(), ref_sp.sum(1).mean())
        prof_d, prof_r = dev_sp.mean(0), ref_sp.mean(0)
        assert np.abs(prof_d - prof_r).sum() / prof_r.sum() < 0.25, ("mean setpoint profile", np.abs(prof_d - prof_r).sum() / prof_r.sum())
        assert _ks(dev_sp[dev_sp > 0], ref_sp[ref_sp > 0]) < 2.0 * ks_bar, ("setpoint values", _ks(dev_sp[dev_sp > 0], ref_sp[ref_sp > 0]))
        assert abs((dev_sp > 0).mean() - (ref_sp > 0).mean()) < 0.05, "share of steps with a setpoint"
    else:
        assert np.array_equal(dev_sp, ref_sp), "setpoints stay the bank's when the config does not derive them"
    # the same seed reproduces the same sessions (and setpoints)
    eng.resample_sessions(seed=1234)
    again = eng.read_sessions(0)
    assert all(np.array_equal(again[k], first[k], equal_nan=True) for k in first)
    if n_rounds == 1:
        assert np.array_equal(eng.read_setpoints(0), dev_sp[0])
    # ---- (3) an episode on the sampled sessions == the oracle on the sessions read back
    if steps_vs_oracle:
        E = S
        eng.reset()
        scns = []
        for s in range(E):
            d = eng.read_sessions(s)
            base = pack.scenarios[s]
            m = d["model"]
            sess = dict(loc=(np.searchsorted(topo.cs_port_off, d["port"], side="right") - 1).astype(np.int32),
                        t_arr=d["t_arr"], t_dep=d["t_dep"], ev_phases=tab.model_phases[m].astype(np.int32),
                        lut=tab.model_lut[m].astype(np.int32), cap0=d["cap0"], B=tab.model_B[m], pmax_ac=tab.model_pmax_ac[m],
                        pmin_ac=tab.model_pmin_ac[m], pmax_dis=tab.model_pmax_dis[m], pmin_dis=tab.model_pmin_dis[m],
                        bmin=np.full(len(m), tab.min_battery_capacity),
                        bmin_em=np.where(tab.min_emergency_battery_capacity > tab.model_B[m], 0.7 * tab.model_B[m], tab.min_emergency_battery_capacity),
                        desired=tab.desired_frac * tab.model_B[m],
                        ts=np.where(np.isnan(d["ts"]), tab.homog_ts, d["ts"]), mult=np.full(len(m), tab.ts_multiplier),
                        eta_c=np.where(np.isnan(d["eta_c"]), 1.0, d["eta_c"]), eta_d=np.where(np.isnan(d["eta_d"]), 1.0, d["eta_d"]))
            setp = np.array(base.setpoint, dtype=np.float64)
            setp[:T] = eng.read_setpoints(s)
            scns.append(Scenario(charge_price=base.charge_price, discharge_price=base.discharge_price, setpoint=setp,
                                 tr_infl=base.tr_infl, tr_solar=base.tr_solar, tr_max_power=base.tr_max_power,
                                 tr_min_power=base.tr_min_power, tr_load_fc=base.tr_load_fc, tr_pv_fc=base.tr_pv_fc,
                                 dr_start=base.dr_start, dr_end=base.dr_end, dr_cap=base.dr_cap, dr_count=base.dr_count,
                                 sessions=sess, luts_c=np.asarray(tab.luts).reshape(-1, 101), luts_d=np.asarray(tab.luts).reshape(-1, 101)).normalise())
        orc = OracleBatch(topo, scns, reward=reward, state=state)
        orc.reset()
        rng = np.random.default_rng(3)
        low = -1.0 if topo.v2g_enabled else 0.0
        caps = eng.state()["port_cap"] if hasattr(eng, "state") else None
        for t in range(steps_vs_oracle):
            a = rng.uniform(low, 1.0, (E, P))
            out = eng.step(np.ascontiguousarray(a)) if caps is not None else eng.step_np(a)
            orc.step(a)
            occ = orc.arr["port_session"] >= 0
            c = caps if caps is not None else eng.state_tensors()["port_cap"].cpu().numpy()
            assert np.array_equal(c[occ], orc.arr["port_cap"][occ]), (t, "battery level")
            r = out["reward"] if caps is not None else out["reward"].cpu().numpy()
            assert np.all(np.abs(r - orc.reward) <= 1e-9 + 1e-9 * np.abs(orc.reward)), (t, "reward")
            o = out["obs"] if caps is not None else out["obs"].cpu().numpy()
            assert np.allclose(o, orc.o["obs"][:, :eng.D], rtol=1e-5, atol=1e-5), (t, "obs")
    eng.close()


@pytest.mark.parametrize("name,reward,state", CASES[:2])
def test_device_sampler_on_emulator(name, reward, state, monkeypatch):
    import emu_engine
    emu_engine.build()
    monkeypatch.setenv("EV2B_KERNEL", "evlist")

    def make(topo, E, rw, st):
        return emu_engine.EmuEngine(topo, E, reward=rw, state=st, outputs=("reward", "status", "obs"))
    # 64 scenarios x 1 round: ~9k sessions (c3) / ~1.5k (c2) against as many of the reference: KS noise ~ 1.36 * sqrt(2 / n)
    _check_sample(make, name, reward, state, n_rounds=1, ks_bar=0.06 if "c3" in name else 0.09, steps_vs_oracle=40 if "c2" in name else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,reward,state", CASES)
def test_device_sampler_on_gpu(name, reward, state, monkeypatch):
    import torch
    from ev2gym_b200.engine import BatchedEngine
    monkeypatch.setenv("EV2B_KERNEL", "evlist")

    class Eng(BatchedEngine):
        def step_np(self, a):
            return self.step(torch.tensor(a, device="cuda"))

    def make(topo, E, rw, st):
        return Eng(topo, E, reward=rw, state=st, outputs=("reward", "status", "obs"))
    _check_sample(make, name, reward, state, n_rounds=8, ks_bar=0.05 if "c3" in name else 0.07, steps_vs_oracle=112)


# ---- the same seed gives the same sessions on the emulated and on the real build ---------------------------------------
GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLER_GOLDEN = [("c2_publicpst_c25", "SquaredTrackingErrorReward", "PublicPST"),
                  ("c3_v2gloads_c100n2tr5", "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads")]


def _check_against_sampler_golden(make_engine, name, reward, state):
    """tests/golden/sampler_<bank>.npz (tools/make_sampler_golden.py: the emulated kernels, which
    tests/test_spawner_reference.py / test_setpoints_reference.py pin to the unmodified reference functions)."""
    pack, tab = _load(name)
    z = np.load(os.path.join(GOLDEN, "sampler_" + name + ".npz"))
    S = int(z["n_scenarios"][0])
    eng = make_engine(pack.topo, S, reward, state)
    eng.set_spawn_tables(tab)
    eng.load_scenarios(pack.scenarios[:S])
    eng.resample_sessions(seed=int(z["seed"][0]))
    n = 0
    for s in range(S):
        d = eng.read_sessions(s)
        for k in ("port", "t_arr", "t_dep", "model"):
            assert np.array_equal(d[k], z[f"s{s}_{k}"]), (s, k)
        assert np.allclose(d["cap0"], z[f"s{s}_cap0"], rtol=1e-9, atol=1e-9), s
        for k in ("ts", "eta_c", "eta_d"):
            assert np.array_equal(d[k], z[f"s{s}_{k}"], equal_nan=True), (s, k)
        sp = eng.read_setpoints(s)
        assert np.allclose(sp, z[f"s{s}_setpoint"], rtol=1e-9, atol=1e-9), (s, "setpoints")
        n += len(d["port"])
    assert n > 200
    eng.close()


@pytest.mark.parametrize("name,reward,state", SAMPLER_GOLDEN)
def test_emulated_sampler_equals_the_golden_sessions(name, reward, state, monkeypatch):
    """(guards the fixture: it must be regenerated when the sampler's keying or arithmetic changes)"""
    import emu_engine
    emu_engine.build()
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    _check_against_sampler_golden(lambda topo, E, rw, st: emu_engine.EmuEngine(topo, E, reward=rw, state=st, outputs=("reward",)),
                                  name, reward, state)


@pytest.mark.gpu
@pytest.mark.parametrize("name,reward,state", SAMPLER_GOLDEN)
def test_gpu_sampler_equals_the_golden_sessions(name, reward, state, monkeypatch):
    from ev2gym_b200.engine import BatchedEngine
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    _check_against_sampler_golden(lambda topo, E, rw, st: BatchedEngine(topo, E, reward=rw, state=st, outputs=("reward",)),
                                  name, reward, state)
