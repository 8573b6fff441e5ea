"""Export a *reference* `EV2Gym` object (after `reset()`) into `Topology` / `Scenario`.

This is the drop-in "scenario source" for users who have the reference installed: the
reference's own host-side generators (loaders.py / utils.py, out of scope per SURVEY.md
section 2 rows 5 and 7) keep producing the episode, and the B200 engine consumes the
exported tensors.  This module never imports the reference; it only reads attributes of the
object it is handed (duck typing), so it also works on unpickled replay scenarios.

Attribute provenance (paths relative to /root/reference):
  env.charging_stations[*]    ev2gym/models/ev_charger.py:41-75
  env.transformers[*]         ev2gym/models/transformer.py:15-78
  env.EVs_profiles[*]         ev2gym/models/ev.py:45-113, built by ev2gym/utilities/utils.py:298-345
  env.charge_prices etc.      ev2gym/models/ev2gym_env.py:293-296
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

from .scenario import LUT_LEN, SESSION_F64_FIELDS, SESSION_INT_FIELDS, Scenario, SpawnTables, Topology


def topology_from_env(env) -> Topology:
    cs = env.charging_stations
    cfg = env.config
    return Topology(
        cs_n_ports=np.array([c.n_ports for c in cs]),
        cs_tr=np.array([c.connected_transformer for c in cs]),
        cs_imax=np.array([c.max_charge_current for c in cs], dtype=np.float64),
        cs_imin=np.array([c.min_charge_current for c in cs], dtype=np.float64),
        cs_imax_dis=np.array([c.max_discharge_current for c in cs], dtype=np.float64),
        cs_imin_dis=np.array([c.min_discharge_current for c in cs], dtype=np.float64),
        cs_voltage=np.array([c.voltage for c in cs], dtype=np.float64),
        cs_phases=np.array([c.phases for c in cs]),
        n_transformers=len(env.transformers),
        tr_voltage=float(env.transformers[0].voltage) if len(env.transformers) else
        float(cfg["charging_station"]["voltage"] * math.sqrt(cfg["charging_station"]["phases"])),
        timescale=int(env.timescale),
        sim_length=int(env.simulation_length),
        dr_steps_ahead=int(cfg["demand_response"]["notification_of_event_minutes"] // env.timescale),
        v2g_enabled=bool(cfg["v2g_enabled"]),
        **_grid_static(env),
    )


def _grid_static(env) -> dict:
    """K, L of the Laurent power flow (grid_tensor.py:110-118) when `simulate_grid` is on."""
    grid = getattr(env, "grid", None)
    if not getattr(env, "simulate_grid", False) or grid is None:
        return {}
    net = grid.net
    return dict(grid_K=np.asarray(net._K_, dtype=np.complex128), grid_L=np.asarray(net._L_, dtype=np.complex128).reshape(-1),
                grid_s_base=float(net.s_base))


def _grid_series(env, T: int):
    """Base bus powers of steps 0..T exactly as PowerGrid.reset/step derive them (grid.py:109-118, 131-139)."""
    grid = env.grid
    nb = grid.node_num
    act = np.zeros((T + 1, nb - 1))
    rea = np.zeros((T + 1, nb - 1))
    # step 0: PowerGrid.reset's `active_power` is a VIEW of load_data[0] and is modified in place (`-= pv`,
    # grid.py:109-116), so row 0 of load_data no longer holds the raw load: take what reset() left behind.
    act[0] = np.asarray(grid.active_power, dtype=np.float64).reshape(-1)
    rea[0] = np.asarray(grid.reactive_power, dtype=np.float64).reshape(-1)
    for t in range(1, T + 1):
        a = np.array(grid.load_data[t, 1:nb], dtype=np.float64).reshape(1, -1)
        r = (a * grid.net.pf).round(1)
        a = a - np.asarray(grid.pv_data[t, 1:nb], dtype=np.float64).reshape(1, -1)
        act[t], rea[t] = a[0], r[0]
    return act, rea


class _ReplayAsEnv:
    """Presents an `EvCityReplay` (ev2gym/models/replay.py:10-182) under the attribute names `topology_from_env` /
    `scenario_from_env` read -- the same objects `EV2Gym(load_from_replay_path=...)` hands to its loaders
    (loaders.py:97-98, 235-236, 307-308, 389, 400-401)."""

    def __init__(self, replay, config: dict):
        if getattr(replay, "simulate_grid", False):
            raise NotImplementedError("grid replays are not imported: PowerGrid.reset modifies load_data[0] in place "
                                      "(grid.py:109-116), so a saved grid replay no longer holds the raw step-0 loads")
        self.charging_stations, self.transformers = replay.charging_stations, replay.transformers
        self.EVs_profiles = replay.EVs
        self.charge_prices, self.discharge_prices = replay.charge_prices, replay.discharge_prices
        self.power_setpoints = replay.power_setpoints
        self.simulation_length, self.timescale = replay.sim_length, replay.timescale
        self.sim_starting_date = self.sim_date = replay.sim_date
        self.simulate_grid, self.grid = False, None
        self.config = config
        self.seed = None


def topology_from_replay(replay, config: dict) -> Topology:
    """Topology of a saved reference episode.  `config` is the YAML dict the reference would be constructed with
    (ev2gym_env.py:64-65): the replay does not store `v2g_enabled` or the demand-response notification time."""
    return topology_from_env(_ReplayAsEnv(replay, config))


def scenario_from_replay(replay, config: dict) -> Scenario:
    """Scenario of a saved reference episode: what `EV2Gym(load_from_replay_path=...)` would re-run."""
    return scenario_from_env(_ReplayAsEnv(replay, config))


def _lut_row(d: dict) -> Tuple[float, ...]:
    """`dict.get(k, 1)` for k = 0..100, the only keys `np.round(amps)` can hit within LUT_LEN
    (ev.py:287-288; the dict is dense over 0..100 after utils.py:282-288)."""
    bad = [k for k in d if not (0 <= k < LUT_LEN) or k != int(k)]
    if bad:
        raise ValueError(f"efficiency table has keys outside 0..{LUT_LEN - 1}: {bad[:5]}")
    return tuple(float(d.get(k, 1)) for k in range(LUT_LEN))


def scenario_from_env(env) -> Scenario:
    """Snapshot everything `reset()` sampled.  Call right after `reset()` (before any `step`)."""
    T = int(env.simulation_length)
    trs = env.transformers
    Tr = len(trs)
    cp, dp = np.asarray(env.charge_prices, dtype=np.float64), np.asarray(env.discharge_prices, dtype=np.float64)
    if not (np.all(cp == cp[0:1]) and np.all(dp == dp[0:1])):
        raise ValueError("per-charger prices differ; the engine stores one price row per env (loaders.py:423-424)")

    def series(name):
        return np.stack([np.asarray(getattr(tr, name), dtype=np.float64)[:T] for tr in trs]) if Tr else \
            np.zeros((0, T))

    ndr = max([len(tr.dr_events) for tr in trs] + [1])
    dr_start = np.zeros((Tr, ndr), dtype=np.int32)
    dr_end = np.zeros((Tr, ndr), dtype=np.int32)
    dr_cap = np.zeros((Tr, ndr))
    dr_count = np.zeros(Tr, dtype=np.int32)
    for i, tr in enumerate(trs):
        dr_count[i] = len(tr.dr_events)
        for j, ev in enumerate(tr.dr_events):
            dr_start[i, j], dr_end[i, j], dr_cap[i, j] = ev["event_start_step"], ev["event_end_step"], \
                float(ev["capacity_percentage"])

    sess = {k: [] for k in SESSION_INT_FIELDS + SESSION_F64_FIELDS}
    luts_c, luts_d, lut_index = [], [], {}
    for ev in env.EVs_profiles:
        ce, de = ev.charge_efficiency, ev.discharge_efficiency
        if isinstance(ce, dict):
            # ev.py:375 tests the *charge* efficiency's type before using the discharge dict
            key = (_lut_row(ce), _lut_row(de))
            if key not in lut_index:
                lut_index[key] = len(luts_c)
                luts_c.append(key[0])
                luts_d.append(key[1])
            lut, eta_c, eta_d = lut_index[key], math.nan, math.nan
        else:
            lut, eta_c, eta_d = -1, float(ce), float(de)
        vals = dict(loc=ev.location, t_arr=ev.time_of_arrival, t_dep=ev.time_of_departure,
                    ev_phases=ev.ev_phases, lut=lut, cap0=ev.battery_capacity_at_arrival,
                    B=ev.battery_capacity, pmax_ac=ev.max_ac_charge_power, pmin_ac=ev.min_ac_charge_power,
                    pmax_dis=ev.max_discharge_power, pmin_dis=ev.min_discharge_power,
                    bmin=ev.min_battery_capacity, bmin_em=ev.min_emergency_battery_capacity,
                    desired=ev.desired_capacity, ts=ev.transition_soc, mult=ev.transition_soc_multiplier,
                    eta_c=eta_c, eta_d=eta_d)
        for k, v in vals.items():
            sess[k].append(v)
    sessions = {k: np.array(v, dtype=np.int32 if k in SESSION_INT_FIELDS else np.float64) for k, v in sess.items()}

    import datetime
    start = env.sim_starting_date if hasattr(env, "sim_starting_date") else env.sim_date
    dates = [start + datetime.timedelta(minutes=int(env.timescale) * k) for k in range(T + 1)]
    date_feat = np.array([[d.weekday() / 7, math.sin(d.hour / 24 * 2 * math.pi), math.cos(d.hour / 24 * 2 * math.pi)]
                          for d in dates])                                    # state.py:221-225
    grid_kw = {}
    if getattr(env, "simulate_grid", False) and getattr(env, "grid", None) is not None:
        ga, gr = _grid_series(env, T)
        grid_kw = dict(grid_active=ga, grid_reactive=gr)
    return Scenario(
        date_feat=date_feat, **grid_kw,
        charge_price=cp[0].copy(), discharge_price=dp[0].copy(),
        setpoint=np.asarray(env.power_setpoints, dtype=np.float64)[:T].copy(),
        tr_infl=series("inflexible_load"), tr_solar=series("solar_power"),
        tr_max_power=series("max_power"), tr_min_power=series("min_power"),
        tr_load_fc=series("inflexible_load_forecast"), tr_pv_fc=series("pv_generation_forecast"),
        dr_start=dr_start, dr_end=dr_end, dr_cap=dr_cap, dr_count=dr_count,
        sessions=sessions,
        luts_c=np.array(luts_c, dtype=np.float64).reshape(-1, LUT_LEN),
        luts_d=np.array(luts_d, dtype=np.float64).reshape(-1, LUT_LEN),
        meta={"sim_date": str(getattr(env, "sim_date", "")), "seed": getattr(env, "seed", None)},
    ).normalise()


def spawn_tables_from_env(env, starts=None) -> SpawnTables:
    """The tables and scalars EV_spawner / spawn_single_EV read from `env` (utils.py:477-557, 177-345), as arrays.
    `starts`: optional list of sim_date values (one per scenario of the bank the tables go with)."""
    cfg, scenario = env.config, env.scenario
    half_hours = [f"{h:02d}:{m:02d}" for h in range(24) for m in (0, 30)]

    def by_arrival(df):
        col = df.set_index("Arrival Time")[scenario]
        return np.array([float(col.get(k, 0.0)) for k in half_hours], dtype=np.float64)

    luts, model_lut = [], []
    if cfg["heterogeneous_ev_specs"]:
        names = list(env.ev_specs.keys())
        prob = np.asarray(env.normalized_ev_registrations, dtype=np.float64)
        B = [env.ev_specs[n]["battery_capacity"] for n in names]
        pac = [env.ev_specs[n]["max_ac_charge_power"] for n in names]
        pdis = [-env.ev_specs[n]["max_ac_discharge_power"] for n in names]
        pmin_ac, pmin_dis, phases = [0.0] * len(names), [0.0] * len(names), [3] * len(names)     # EV() defaults / ev_phases=3
        for n in names:
            spec = env.ev_specs[n]
            if "3ph_ch_efficiency" in spec:                                    # utils.py:273-290
                eff = dict(zip(spec["ch_current"], spec["3ph_ch_efficiency"]))
                for i in range(0, 101):
                    if i not in eff or eff[i] == 0:
                        nz = [k for k, v in eff.items() if v != 0]
                        if nz:
                            eff[i] = eff[min(nz, key=lambda x: abs(x - i))]
                row = [float(eff.get(i, 1.0)) for i in range(LUT_LEN)]
                if row not in luts:
                    luts.append(row)
                model_lut.append(luts.index(row))
            else:
                model_lut.append(-1)
    else:
        ev = cfg["ev"]
        prob, B, pac, pdis = np.ones(1), [ev["battery_capacity"]], [ev["max_ac_charge_power"]], [ev["max_discharge_power"]]
        pmin_ac, pmin_dis, phases = [ev["min_ac_charge_power"]], [ev["min_discharge_power"]], [ev["ev_phases"]]
        model_lut = [-1]
    start = np.array([[d.weekday(), d.hour, d.minute] for d in (starts or [])], dtype=np.int32).reshape(-1, 3)
    ev = cfg["ev"]
    return SpawnTables(
        workplace=int(scenario == "workplace"),
        arrival_week=np.asarray(env.df_arrival_week[scenario].to_numpy()[:96], dtype=np.float64),
        arrival_weekend=(np.asarray(env.df_arrival_weekend[scenario].to_numpy()[:96], dtype=np.float64)
                         if scenario in env.df_arrival_weekend.columns else np.zeros(96)),     # (no weekend column for "workplace": closed)
        req_energy_mean=by_arrival(env.df_req_energy), stay_mean=by_arrival(env.df_time_of_stay_vs_arrival),
        spawn_multiplier=float(cfg["spawn_multiplier"]), min_stay_steps=int(ev["min_time_of_stay"] // env.timescale),
        desired_frac=float(ev["desired_capacity"]), min_battery_capacity=float(ev["min_battery_capacity"]),
        min_emergency_battery_capacity=float(ev["min_emergency_battery_capacity"]),
        ts_multiplier=float(ev.get("transition_soc_multiplier", 1)),
        empty_ports_at_end=int(bool(env.empty_ports_at_end_of_simulation)),
        heterogeneous=int(bool(cfg["heterogeneous_ev_specs"])),
        power_setpoint_enabled=int(bool(cfg.get("power_setpoint_enabled", False))),
        power_setpoint_flexibility=float(cfg.get("power_setpoint_flexiblity", 0.0)),
        model_prob=prob, model_B=np.array(B, dtype=np.float64), model_pmax_ac=np.array(pac, dtype=np.float64),
        model_pmax_dis=np.array(pdis, dtype=np.float64), model_pmin_ac=np.array(pmin_ac, dtype=np.float64),
        model_pmin_dis=np.array(pmin_dis, dtype=np.float64), model_phases=np.array(phases, dtype=np.int32),
        model_lut=np.array(model_lut, dtype=np.int32), luts=np.array(luts, dtype=np.float64).reshape(-1, LUT_LEN),
        homog_ts=float(ev.get("transition_soc", 1)), homog_eta_c=float(ev.get("charge_efficiency", 1)),
        homog_eta_d=float(ev.get("discharge_efficiency", 1)), start=start)
