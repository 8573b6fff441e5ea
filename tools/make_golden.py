#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING the unmodified Python reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tools/make_golden.py            # all cases -> tests/golden/*.npz
    python tools/make_golden.py --packs    # scenario banks for bench.py -> ev2gym_b200/data/*.npz

The reference is imported read-only from /root/reference through the stub packages in
oracle/refshim (gymnasium / matplotlib / pandapower / multicopula are absent here, SURVEY.md
section 8c).  Nothing from the reference is copied: fixtures hold only *inputs* (the scenario
`reset()` sampled, the action sequence) and *outputs* (what `step()` returned / left in the
env), as arrays.

Each fixture = one episode: config (a shipped YAML + size overrides), seed, agent.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True
os.chdir(REF)  # shipped configs use ./ev2gym/data/... relative paths

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

import yaml  # noqa: E402
from ev2gym.models.ev2gym_env import EV2Gym  # noqa: E402
from ev2gym.rl_agent import reward as ref_reward  # noqa: E402
from ev2gym.rl_agent import state as ref_state  # noqa: E402
from ev2gym.baselines import heuristics as ref_agents  # noqa: E402

from ev2gym_b200.reference_export import scenario_from_env, topology_from_env  # noqa: E402
from ev2gym_b200.scenario import ScenarioPack  # noqa: E402

import ev2gym.models.data_augment as _da  # noqa: E402


def _synthetic_bus_loads(self, n_buses, n_steps, start_day, start_step):
    """Stand-in for the un-vendored copula sampler (`multicopula`, data_augment.py:68-92), which only feeds
    INPUTS (bus load shapes) to the grid path: a deterministic daily sinusoid + noise per bus."""
    rng = np.random.default_rng(1000 * start_day + start_step)
    x = np.linspace(0, 2 * np.pi * n_steps / 96, n_steps)[:, None]
    return 0.5 + 0.3 * np.sin(x + rng.uniform(0, 6, (1, n_buses))) + 0.1 * rng.random((n_steps, n_buses))


_da.DataGenerator.sample_data = _synthetic_bus_loads

PAIRS = {  # config -> (state, reward)   train_stable_baselines.py:39-51
    "V2Ggrid": ("V2G_grid_state", "V2G_grid_full_reward"),
    "PublicPST": ("PublicPST", "SquaredTrackingErrorReward"),
    "V2GProfitMax": ("V2G_profit_max", "profit_maximization"),
    "V2GProfitPlusLoads": ("V2G_profit_max_loads", "ProfitMax_TrPenalty_UserIncentives"),
}


def make_config(base: str, overrides: dict) -> str:
    cfg = yaml.safe_load(open(f"{REF}/ev2gym/example_config_files/{base}.yaml"))

    def merge(d, o):
        for k, v in o.items():
            if isinstance(v, dict):
                merge(d[k], v)
            else:
                d[k] = v
    merge(cfg, overrides)
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    return f.name


def make_env(base, overrides, seed, state=None, reward=None):
    st, rw = PAIRS[base]
    st, rw = state or st, reward or rw
    path = make_config(base, overrides)
    env = EV2Gym(config_file=path, seed=seed, state_function=getattr(ref_state, st),
                 reward_function=getattr(ref_reward, rw))
    os.unlink(path)
    return env, st, rw


def make_actions(kind: str, env, rng) -> np.ndarray:
    P = env.number_of_ports
    low = -1.0 if env.config["v2g_enabled"] else 0.0
    if kind == "afap":
        return np.ones(P)
    if kind == "zeros":
        return np.zeros(P)
    if kind == "discharge":
        return -np.ones(P)
    if kind == "uniform":
        return rng.uniform(low, 1.0, P)
    if kind == "mixed":   # exact zeros, saturations, tiny values under the min-current gate, |sum|>1 on multi-port
        a = rng.uniform(low, 1.0, P)
        r = rng.random(P)
        a[r < 0.15] = 0.0
        a[(r >= 0.15) & (r < 0.25)] = 1.0
        a[(r >= 0.25) & (r < 0.32)] = low
        a[(r >= 0.32) & (r < 0.40)] *= 1e-3
        return a
    raise ValueError(kind)


def record_episode(base, overrides, seed, agent, state=None, reward=None):
    overrides = dict(overrides)
    reward = overrides.pop("_reward", reward)
    env, st, rw = make_env(base, overrides, seed, state, reward)
    obs0, _ = env.reset(seed=seed)
    topo, scn = topology_from_env(env), scenario_from_env(env)
    T, P, C, Tr = env.simulation_length, env.number_of_ports, env.cs, len(env.transformers)
    rng = np.random.default_rng(seed + 7919)
    rr = ref_agents.RoundRobin(env) if agent == "roundrobin" else \
        ref_agents.ChargeAsLateAsPossible() if agent == "calap" else None

    captured = {}
    inner = env.reward_function

    def spy(e, total_costs, sat_list, invalid):
        captured.update(total_costs=total_costs, sat=list(sat_list), invalid=invalid)
        return inner(e, total_costs, sat_list, invalid)
    env.reward_function = spy

    rec = {k: [] for k in ("actions", "actions_eff", "obs", "reward", "done", "cs_power", "cs_current", "tr_power",
                           "tr_amps", "tr_overload", "cap", "port_t_arr", "energy_exch", "total_costs", "invalid",
                           "n_departed", "sat_sum", "action_mask")}
    stats = None
    for t in range(T):
        a = rr.get_action(env) if rr is not None else make_actions(agent, env, rng)
        a = np.asarray(a, dtype=np.float64)
        rec["actions"].append(a.copy())
        obs, r, done, trunc, info = env.step(a)       # mutates `a` in place (ev_charger.py:137-140)
        rec["actions_eff"].append(a.copy())
        rec["obs"].append(np.asarray(obs, dtype=np.float64))
        rec["reward"].append(float(r))
        rec["done"].append(bool(done))
        rec["cs_power"].append(env.cs_power[:, t].copy())
        rec["cs_current"].append(env.cs_current[:, t].copy())
        rec["tr_power"].append(np.array([tr.current_power for tr in env.transformers], dtype=np.float64))
        rec["tr_amps"].append(np.array([tr.current_amps for tr in env.transformers], dtype=np.float64))
        rec["tr_overload"].append(env.tr_overload[:, t].copy())
        cap = np.full(P, np.nan)
        tarr = np.full(P, -1, dtype=np.int32)
        exch = np.zeros(P)
        p = 0
        for cs in env.charging_stations:
            for ev in cs.evs_connected:
                if ev is not None:
                    cap[p], tarr[p], exch[p] = ev.current_capacity, ev.time_of_arrival, ev.total_energy_exchanged
                p += 1
        rec["cap"].append(cap)
        rec["port_t_arr"].append(tarr)
        rec["energy_exch"].append(exch)
        rec["total_costs"].append(float(captured["total_costs"]))
        rec["invalid"].append(int(captured["invalid"]))
        rec["n_departed"].append(len(captured["sat"]))
        rec["sat_sum"].append(float(np.sum(captured["sat"])) if captured["sat"] else 0.0)
        rec["action_mask"].append(np.asarray(info["action_mask"], dtype=np.float64))
        if done:
            stats = info
    out = {k: np.array(v) for k, v in rec.items()}
    out["obs0"] = np.asarray(obs0, dtype=np.float64)
    out["usage"] = env.current_power_usage.copy()
    out["potential"] = env.charge_power_potential.copy()
    out["total_reward"] = np.array(env.total_reward)
    for k in ("total_ev_served", "total_profits", "total_energy_charged", "total_energy_discharged",
              "average_user_satisfaction", "total_transformer_overload", "tracking_error",
              "energy_tracking_error", "power_tracker_violation", "energy_user_satisfaction",
              "std_energy_user_satisfaction", "min_energy_user_satisfaction",
              "total_steps_min_emergency_battery_capacity_violation", "battery_degradation",
              "battery_degradation_calendar", "battery_degradation_cycling", "total_reward"):
        out["stat_" + k] = np.array(float(stats[k]))
    out["afap"] = np.array([ev.max_energy_AFAP for ev in env.EVs], dtype=np.float64)   # ev.py:407-440, spawn order
    if env.simulate_grid:
        out["node_voltage"] = env.node_voltage.copy()                   # [nb, T]
        out["node_active_power"] = env.node_active_power.copy()
        out["node_reactive_power"] = env.node_reactive_power.copy()
        out["node_ev_power"] = env.node_ev_power.copy()
    out["state_fn"], out["reward_fn"] = np.array(st), np.array(rw)
    out["seed"], out["agent"], out["base"] = np.array(seed), np.array(agent), np.array(base)
    return topo, scn, out


HOMOG = {"heterogeneous_ev_specs": False}
CASES = [
    # name, base config, overrides, seed, agent
    ("c1_afap_s42", "V2GProfitPlusLoads", {"number_of_charging_stations": 10}, 42, "afap"),          # KA7
    ("c1_uniform_s7", "V2GProfitPlusLoads", {"number_of_charging_stations": 10}, 7, "uniform"),
    ("pst25_uniform_s3", "PublicPST", {"number_of_charging_stations": 25}, 3, "uniform"),           # c2 shape
    ("pst25_roundrobin_s5", "PublicPST", {"number_of_charging_stations": 25}, 5, "roundrobin"),
    ("loads_c20n2tr3_mixed_s11", "V2GProfitPlusLoads",
     {"number_of_charging_stations": 20, "number_of_ports_per_cs": 2, "number_of_transformers": 3}, 11, "mixed"),
    ("loads_c100n2tr5_uniform_s1", "V2GProfitPlusLoads",                                              # c3 shape
     {"number_of_charging_stations": 100, "number_of_ports_per_cs": 2, "number_of_transformers": 5}, 1, "uniform"),
    ("loads_c12n3tr2_discharge_s2", "V2GProfitPlusLoads",
     {"number_of_charging_stations": 12, "number_of_ports_per_cs": 3, "number_of_transformers": 2}, 2, "discharge"),
    ("profitmax_c25_mixed_s9", "V2GProfitMax", {"number_of_charging_stations": 25}, 9, "mixed"),      # c4 family
    ("homog_ts1_afap_s4", "V2GProfitPlusLoads", {"number_of_charging_stations": 8, **HOMOG}, 4, "afap"),  # ceil lattice
    ("homog_ts08_mixed_s6", "V2GProfitPlusLoads",
     {"number_of_charging_stations": 8, **HOMOG, "ev": {"transition_soc": 0.8, "charge_efficiency": 0.93,
                                                       "discharge_efficiency": 0.91, "min_ac_charge_power": 1.5}},
     6, "mixed"),
    ("mincur_c6n2_mixed_s8", "PublicPST",
     {"number_of_charging_stations": 6, "number_of_ports_per_cs": 2, "v2g_enabled": True,
      "charging_station": {"min_charge_current": 6, "max_charge_current": 32, "max_discharge_current": -32,
                           "min_discharge_current": -6}}, 8, "mixed"),
    ("grid_c40_uniform_s3", "V2Ggrid", {"number_of_charging_stations": 40}, 3, "uniform"),          # Laurent power flow
    ("grid_c20n2_mixed_s5", "V2Ggrid", {"number_of_charging_stations": 20, "number_of_ports_per_cs": 2,
                                       "_reward": "V2G_grid_simple_reward"}, 5, "mixed"),
    ("pst12n2_roundrobin_s12", "PublicPST", {"number_of_charging_stations": 12, "number_of_ports_per_cs": 2}, 12,
     "roundrobin"),                                                                                  # fractional last EV, 1/n_ports
    ("pst25_calap_s6", "PublicPST", {"number_of_charging_stations": 25}, 6, "calap"),
    ("loads_c10n2_calap_s13", "V2GProfitPlusLoads",
     {"number_of_charging_stations": 10, "number_of_ports_per_cs": 2, "number_of_transformers": 2}, 13, "calap"),
    # the remaining stock reward functions (reward.py), one episode each
    ("pst25_sqtr_s14", "PublicPST", {"number_of_charging_stations": 25, "_reward": "SqTrError_TrPenalty_UserIncentives"},
     14, "uniform"),
    ("pst10_simple_s15", "PublicPST", {"number_of_charging_stations": 10, "_reward": "SimpleReward"}, 15, "uniform"),
    ("pst10_mintracker_s16", "PublicPST",
     {"number_of_charging_stations": 10, "_reward": "MinimizeTrackerSurplusWithChargeRewards"}, 16, "uniform"),
    ("profitmax_c12_v2gprofitmax_s17", "V2GProfitMax", {"number_of_charging_stations": 12, "_reward": "V2G_profitmax"},
     17, "mixed"),
    ("profitmax_c8_costs_s18", "V2GProfitMax", {"number_of_charging_stations": 8, "_reward": "V2G_costs_simple"}, 18, "mixed"),
    ("loads_c12n2tr2_v2_s19", "V2GProfitPlusLoads",
     {"number_of_charging_stations": 12, "number_of_ports_per_cs": 2, "number_of_transformers": 2,
      "_reward": "V2G_profitmaxV2"}, 19, "mixed"),
    ("grid_c16_v2_s20", "V2Ggrid", {"number_of_charging_stations": 16, "_reward": "Grid_V2G_profitmaxV2"}, 20, "mixed"),
    ("pst8n2_pstv2_s21", "PublicPST", {"number_of_charging_stations": 8, "number_of_ports_per_cs": 2, "v2g_enabled": True,
                                       "_reward": "pst_V2G_profitmaxV2"}, 21, "mixed"),
    ("pst4_sqpenalty_s22", "PublicPST", {"number_of_charging_stations": 4,
                                         "_reward": "SquaredTrackingErrorRewardWithPenalty"}, 22, "mixed"),
    ("ts10_c5_uniform_s10", "V2GProfitMax", {"number_of_charging_stations": 5, "timescale": 10,
                                             "simulation_length": 150}, 10, "uniform"),
]


def replay_check() -> int:
    """Round trip of the reference's replay pickle (ev2gym/models/replay.py): save an episode, then
    (a) let the reference re-load it (`load_from_replay_path`) and export that env, (b) import the pickle directly with
    `scenario_from_replay`; both must give the same tensors, and the oracle fed with (b) must reproduce, bit for bit, the
    rewards of the reference re-running the replay."""
    import pickle
    import shutil
    from ev2gym_b200.reference_export import scenario_from_replay, topology_from_replay
    from oracle.oracle import OracleEnv
    ov = {"number_of_charging_stations": 6, "number_of_ports_per_cs": 2, "number_of_transformers": 2}
    st, rw = PAIRS["V2GProfitPlusLoads"]
    path = make_config("V2GProfitPlusLoads", ov)
    cfg = yaml.safe_load(open(path))
    tmp = tempfile.mkdtemp()
    try:
        env = EV2Gym(config_file=path, seed=21, save_replay=True, replay_save_path=tmp + "/",
                     state_function=getattr(ref_state, st), reward_function=getattr(ref_reward, rw))
        env.reset(seed=21)
        rng = np.random.default_rng(3)
        acts = [rng.uniform(-1, 1, env.number_of_ports) for _ in range(env.simulation_length)]
        for t, a in enumerate(acts):
            if t == env.simulation_length - 1:
                os.chdir(tmp)                    # EvCityReplay.__init__ creates ./replay (replay.py:19-20)
            env.step(a.copy())
        os.chdir(REF)
        files = [f for f in os.listdir(tmp) if f.endswith(".pkl")]
        assert len(files) == 1, files
        rp = os.path.join(tmp, files[0])
        env2 = EV2Gym(config_file=path, seed=5, load_from_replay_path=rp,
                      state_function=getattr(ref_state, st), reward_function=getattr(ref_reward, rw))
        env2.reset()
        topo_a, scn_a = topology_from_env(env2), scenario_from_env(env2)
        replay = pickle.load(open(rp, "rb"))
        topo_b, scn_b = topology_from_replay(replay, cfg), scenario_from_replay(replay, cfg)
        da, db = ScenarioPack(topo_a, [scn_a], "a").to_dict(), ScenarioPack(topo_b, [scn_b], "b").to_dict()
        bad = [k for k in da if k not in ("config_name",) and not k.endswith("meta") and
               not np.array_equal(np.asarray(da[k]), np.asarray(db[k]), equal_nan=np.asarray(da[k]).dtype.kind == "f")]
        assert not bad, bad
        orc = OracleEnv(topo_b, scn_b, reward=rw, state=st)
        orc.reset()
        for t, a in enumerate(acts):
            _, r, done, _, _ = env2.step(a.copy())
            assert orc.step(a)["reward"] == float(r), t
        assert done
        print(f"replay check: {len(da)} tensors equal, {len(acts)} rewards bit-equal, sessions={scn_b.n_sessions}")
        return 0
    finally:
        os.chdir(REF)
        shutil.rmtree(tmp, ignore_errors=True)
        os.unlink(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replay-check", action="store_true", help="round-trip a reference replay pickle (no files written)")
    ap.add_argument("--packs", action="store_true", help="write the scenario banks used by bench.py")
    ap.add_argument("--spawn-tables", action="store_true",
                    help="write the EV-spawner tables of those banks (ev2gym_b200/data/spawn_*.npz) for the device sampler")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if args.replay_check:
        sys.exit(replay_check())
    if args.packs or args.spawn_tables:
        out_dir = os.path.join(REPO, "ev2gym_b200", "data")
        os.makedirs(out_dir, exist_ok=True)
        banks = [
            ("c2_publicpst_c25", "PublicPST", {"number_of_charging_stations": 25}, 64),
            ("c3_v2gloads_c100n2tr5", "V2GProfitPlusLoads",
             {"number_of_charging_stations": 100, "number_of_ports_per_cs": 2, "number_of_transformers": 5}, 64),
            ("c4_v2gprofitmax_c250", "V2GProfitMax", {"number_of_charging_stations": 250}, 32),
        ]
        for name, base, ov, n in banks:
            if args.only and args.only not in name:
                continue
            env, _, _ = make_env(base, ov, 1000)
            scns, starts = [], []
            for i in range(n):      # seed = 1000 + i
                env.reset(seed=1000 + i)
                starts.append(env.sim_date)
                if args.packs:
                    scns.append(scenario_from_env(env))
            if args.packs:
                ScenarioPack(topology_from_env(env), scns, config_name=name).save(os.path.join(out_dir, name + ".npz"))
                print("pack", name, n, "scenarios")
            if args.spawn_tables:
                from ev2gym_b200.reference_export import spawn_tables_from_env
                spawn_tables_from_env(env, starts).save(os.path.join(out_dir, "spawn_" + name + ".npz"))
                print("spawn tables", name, n, "start dates")
        return
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, base, ov, seed, agent in CASES:
        if args.only and args.only not in name:
            continue
        topo, scn, tr = record_episode(base, ov, seed, agent)
        pack = ScenarioPack(topo, [scn], config_name=name)
        pack.save(os.path.join(out_dir, name + ".scenario.npz"))
        np.savez_compressed(os.path.join(out_dir, name + ".trace.npz"), **tr)
        print(f"{name}: P={topo.P} Tr={topo.Tr} sessions={scn.n_sessions} sum_reward={tr['reward'].sum():.12f}")


if __name__ == "__main__":
    main()
