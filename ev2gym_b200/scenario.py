"""Scenario containers: everything `EV2Gym.reset()` samples, as plain fp64/int arrays.

The reference's `step()` consumes no randomness (SURVEY.md "Quick facts"): it is a
deterministic function of the scenario produced by `reset()` and of the action
sequence.  A `Scenario` is that scenario for ONE env replica; a `ScenarioPack` is a
bank of scenarios sharing one charger/transformer `Topology`.

Provenance of every field (reference file:line, relative to /root/reference):
  Topology      <- ev2gym/utilities/loaders.py:299-365 (chargers), :227-296 + :464-500 (cs->tr map),
                   ev2gym/models/transformer.py:39-40 (transformer voltage)
  prices        <- loaders.py:392-461   (identical rows for all chargers -> one row per env)
  setpoints     <- loaders.py:92-103 / ev2gym/utilities/utils.py:664-757
  tr_* series   <- transformer.py:43-78 (after normalisation / DR events / forecasts)
  sessions      <- utils.py:177-345 (`spawn_single_EV`), list order = `env.EVs_profiles`
                   (arrival sorted, utils.py:504-557)
This module is host-side product code (numpy only, no CUDA, no reference import).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

# Session fields (one entry per EV profile, in `EVs_profiles` order).
SESSION_INT_FIELDS = ("loc", "t_arr", "t_dep", "ev_phases", "lut")
SESSION_F64_FIELDS = (
    "cap0",        # battery_capacity_at_arrival            ev.py:76
    "B",           # battery_capacity                       ev.py:79
    "pmax_ac",     # max_ac_charge_power                    ev.py:82
    "pmin_ac",     # min_ac_charge_power                    ev.py:83
    "pmax_dis",    # max_discharge_power (<= 0)             ev.py:84
    "pmin_dis",    # min_discharge_power                    ev.py:85
    "bmin",        # min_battery_capacity                   ev.py:80
    "bmin_em",     # min_emergency_battery_capacity         ev.py:81
    "desired",     # desired_capacity                       ev.py:75
    "ts",          # transition_soc                         ev.py:87
    "mult",        # transition_soc_multiplier              ev.py:88
    "eta_c",       # charge_efficiency (scalar; NaN if LUT) ev.py:91
    "eta_d",       # discharge_efficiency (scalar; NaN if LUT)
)
TR_SERIES = ("tr_infl", "tr_solar", "tr_max_power", "tr_min_power", "tr_load_fc", "tr_pv_fc")
LUT_LEN = 101  # efficiency dict keys 0..100 A (utils.py:282-288)


@dataclass
class Topology:
    """Static charger / transformer layout (identical for every env of a handle)."""
    cs_n_ports: np.ndarray      # [C] int32
    cs_tr: np.ndarray           # [C] int32  connected_transformer
    cs_imax: np.ndarray         # [C] f64    max_charge_current
    cs_imin: np.ndarray         # [C] f64    min_charge_current
    cs_imax_dis: np.ndarray     # [C] f64    max_discharge_current (<= 0)
    cs_imin_dis: np.ndarray     # [C] f64    min_discharge_current
    cs_voltage: np.ndarray      # [C] f64
    cs_phases: np.ndarray       # [C] int32
    n_transformers: int
    tr_voltage: float           # voltage * sqrt(phases) of the config   transformer.py:39-40
    timescale: int              # minutes per step
    sim_length: int             # T
    dr_steps_ahead: int = 0     # notification_of_event_minutes // timescale  transformer.py:69
    v2g_enabled: bool = True
    # distribution grid (simulate_grid: True): Laurent power flow V <- K conj(S/V) + L   grid_tensor.py:110-118
    grid_K: Optional[np.ndarray] = None      # [nb-1, nb-1] complex128  (-Ydd^-1)
    grid_L: Optional[np.ndarray] = None      # [nb-1] complex128        (K @ Yds)
    grid_s_base: float = 1000.0              # kVA

    def __post_init__(self):
        self.cs_n_ports = np.ascontiguousarray(self.cs_n_ports, dtype=np.int32)
        self.cs_tr = np.ascontiguousarray(self.cs_tr, dtype=np.int32)
        self.cs_phases = np.ascontiguousarray(self.cs_phases, dtype=np.int32)
        for k in ("cs_imax", "cs_imin", "cs_imax_dis", "cs_imin_dis", "cs_voltage"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float64))
        if self.grid_K is not None and np.size(self.grid_K) > 0:
            self.grid_K = np.ascontiguousarray(self.grid_K, dtype=np.complex128)
            self.grid_L = np.ascontiguousarray(self.grid_L, dtype=np.complex128).reshape(-1)
        else:
            self.grid_K = self.grid_L = None

    @property
    def n_bus(self) -> int:
        """Buses without the slack (= transformers in grid mode, loaders.py:481); 0 when no grid is simulated."""
        return 0 if self.grid_K is None else int(self.grid_K.shape[0])

    @property
    def C(self) -> int:
        return int(self.cs_n_ports.shape[0])

    @property
    def P(self) -> int:
        return int(self.cs_n_ports.sum())

    @property
    def Tr(self) -> int:
        return int(self.n_transformers)

    @property
    def T(self) -> int:
        return int(self.sim_length)

    @property
    def cs_port_off(self) -> np.ndarray:
        off = np.zeros(self.C + 1, dtype=np.int32)
        np.cumsum(self.cs_n_ports, out=off[1:])
        return off

    @classmethod
    def uniform(cls, C: int, n_ports: int, Tr: int, T: int = 112, timescale: int = 15,
                imax: float = 32.0, imin: float = 0.0, imax_dis: float = -32.0,
                imin_dis: float = 0.0, voltage: float = 400.0, phases: int = 3,
                v2g_enabled: bool = True, dr_steps_ahead: int = 4) -> "Topology":
        """The reference's default layout: charger i -> transformer i mod Tr (loaders.py:495-498)."""
        import math
        return cls(
            cs_n_ports=np.full(C, n_ports), cs_tr=np.arange(C) % Tr,
            cs_imax=np.full(C, imax), cs_imin=np.full(C, imin),
            cs_imax_dis=np.full(C, imax_dis if v2g_enabled else 0.0),
            cs_imin_dis=np.full(C, imin_dis if v2g_enabled else 0.0),
            cs_voltage=np.full(C, voltage), cs_phases=np.full(C, phases),
            n_transformers=Tr, tr_voltage=voltage * math.sqrt(phases), timescale=timescale,
            sim_length=T, dr_steps_ahead=dr_steps_ahead, v2g_enabled=v2g_enabled)

    def to_dict(self) -> Dict[str, np.ndarray]:
        d = {f"topo_{k}": np.asarray(v if v is not None else np.zeros(0)) for k, v in dataclasses.asdict(self).items()}
        return d

    @classmethod
    def from_dict(cls, d) -> "Topology":
        kw = {}
        for f in dataclasses.fields(cls):
            if f"topo_{f.name}" not in d:          # packs written before the field existed: keep the default
                continue
            v = d[f"topo_{f.name}"]
            if f.name in ("n_transformers", "timescale", "sim_length", "dr_steps_ahead"):
                v = int(v)
            elif f.name == "tr_voltage":
                v = float(v)
            elif f.name == "v2g_enabled":
                v = bool(v)
            elif f.name == "grid_s_base":
                v = float(v)
            elif f.name in ("grid_K", "grid_L"):
                v = None if np.size(v) == 0 else v
            kw[f.name] = v
        return cls(**kw)


@dataclass
class Scenario:
    """One env replica's pre-sampled episode (everything `reset()` draws)."""
    charge_price: np.ndarray        # [T] f64  (= -price/1000, loaders.py:439)
    discharge_price: np.ndarray     # [T] f64
    setpoint: np.ndarray            # [T] f64  power_setpoints
    tr_infl: np.ndarray             # [Tr,T] f64 inflexible_load
    tr_solar: np.ndarray            # [Tr,T] f64 solar_power (stored negative, transformer.py:197)
    tr_max_power: np.ndarray        # [Tr,T] f64 (after DR events, transformer.py:118-130)
    tr_min_power: np.ndarray        # [Tr,T] f64
    tr_load_fc: np.ndarray          # [Tr,T] f64 inflexible_load_forecast
    tr_pv_fc: np.ndarray            # [Tr,T] f64 pv_generation_forecast
    dr_start: np.ndarray            # [Tr,NDR] int32 event_start_step
    dr_end: np.ndarray              # [Tr,NDR] int32 event_end_step
    dr_cap: np.ndarray              # [Tr,NDR] f64 capacity_percentage
    dr_count: np.ndarray            # [Tr] int32 number of events
    sessions: Dict[str, np.ndarray] = field(default_factory=dict)  # SESSION_*_FIELDS, each [S]
    luts_c: np.ndarray = field(default_factory=lambda: np.ones((0, LUT_LEN)))  # [L,101] percent
    luts_d: np.ndarray = field(default_factory=lambda: np.ones((0, LUT_LEN)))
    meta: Dict[str, object] = field(default_factory=dict)
    # grid mode only: base bus powers of steps 0..T (grid.py:98-118,131-139) and calendar features of the
    # observation times 0..T (state.py:221-225)
    grid_active: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))    # [T+1, nb-1] kW (load - pv)
    grid_reactive: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))  # [T+1, nb-1] kVAr
    date_feat: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))      # [T+1, 3] weekday/7, sin, cos

    @property
    def n_sessions(self) -> int:
        return int(self.sessions["t_arr"].shape[0])

    def normalise(self) -> "Scenario":
        for k in ("charge_price", "discharge_price", "setpoint", "dr_cap") + TR_SERIES:
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float64))
        for k in ("dr_start", "dr_end", "dr_count"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.int32))
        for k in SESSION_INT_FIELDS:
            self.sessions[k] = np.ascontiguousarray(self.sessions[k], dtype=np.int32)
        for k in SESSION_F64_FIELDS:
            self.sessions[k] = np.ascontiguousarray(self.sessions[k], dtype=np.float64)
        for k in ("grid_active", "grid_reactive", "date_feat"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float64))
        self.luts_c = np.ascontiguousarray(self.luts_c, dtype=np.float64).reshape(-1, LUT_LEN)
        self.luts_d = np.ascontiguousarray(self.luts_d, dtype=np.float64).reshape(-1, LUT_LEN)
        return self


def assign_ports(topo: Topology, t_arr: np.ndarray, t_dep: np.ndarray, loc: np.ndarray) -> np.ndarray:
    """Replay the reference's first-free-port rule on the host.

    `EV_Charger.spawn_ev` puts an arriving EV into `evs_connected.index(None)`
    (ev_charger.py:273), NOT into the port the spawner drew; departures of step t
    (t >= time_of_departure, ev_charger.py:209-224) free their port BEFORE the arrivals
    of t+1 are placed (ev2gym_env.py:363-417).  Sessions must be arrival-sorted with
    t_arr >= 1, which is what `EV_spawner` emits (utils.py:504-557; first arrival is 3).
    Returns the flat port index (charger-major = action order) of every session.
    """
    t_arr = np.asarray(t_arr)
    if t_arr.size and (np.any(np.diff(t_arr) < 0) or t_arr.min() < 1):
        raise ValueError("sessions must be sorted by time_of_arrival and arrive at step >= 1")
    off = topo.cs_port_off
    occupied_until = np.full(topo.P, -1, dtype=np.int64)  # t_dep of the occupant, -1 = free
    port = np.full(t_arr.shape[0], -1, dtype=np.int32)
    for i in range(t_arr.shape[0]):
        t = int(t_arr[i]) - 1                       # the EV is placed at the END of step t
        c = int(loc[i])
        lo, hi = int(off[c]), int(off[c + 1])
        seg = occupied_until[lo:hi]
        seg[(seg >= 0) & (seg <= t)] = -1           # departures of steps <= t already happened
        free = np.nonzero(seg < 0)[0]
        if free.size == 0:
            raise ValueError(f"session {i}: charger {c} has no free port at step {t + 1}")
        seg[free[0]] = int(t_dep[i])
        port[i] = lo + int(free[0])
    return port


@dataclass
class ScenarioPack:
    """A bank of scenarios on one topology, with npz (de)serialisation."""
    topo: Topology
    scenarios: List[Scenario]
    config_name: str = ""

    def __len__(self):
        return len(self.scenarios)

    def save(self, path: str) -> None:
        np.savez_compressed(path, **self.to_dict())

    def to_dict(self) -> Dict[str, np.ndarray]:
        """Flat name -> array form of the whole pack (what `save` writes)."""
        d = dict(self.topo.to_dict())
        d["config_name"] = np.array(self.config_name)
        d["n_scenarios"] = np.array(len(self.scenarios))
        sc = self.scenarios
        for k in ("charge_price", "discharge_price", "setpoint", "dr_start", "dr_end", "dr_cap",
                  "dr_count", "grid_active", "grid_reactive", "date_feat") + TR_SERIES:
            d[k] = np.stack([getattr(s, k) for s in sc])
        s_off = np.zeros(len(sc) + 1, dtype=np.int64)
        l_off = np.zeros(len(sc) + 1, dtype=np.int64)
        for i, s in enumerate(sc):
            s_off[i + 1] = s_off[i] + s.n_sessions
            l_off[i + 1] = l_off[i] + s.luts_c.shape[0]
        d["s_off"], d["l_off"] = s_off, l_off
        for k in SESSION_INT_FIELDS + SESSION_F64_FIELDS:
            d["s_" + k] = np.concatenate([s.sessions[k] for s in sc])
        d["luts_c"] = np.concatenate([s.luts_c for s in sc]).reshape(-1, LUT_LEN)
        d["luts_d"] = np.concatenate([s.luts_d for s in sc]).reshape(-1, LUT_LEN)
        return d

    @classmethod
    def load(cls, path: str) -> "ScenarioPack":
        z = np.load(path, allow_pickle=False)
        topo = Topology.from_dict(z)
        n = int(z["n_scenarios"])
        s_off, l_off = z["s_off"], z["l_off"]
        out = []
        for i in range(n):
            sess = {k: z["s_" + k][s_off[i]:s_off[i + 1]] for k in SESSION_INT_FIELDS + SESSION_F64_FIELDS}
            out.append(Scenario(
                charge_price=z["charge_price"][i], discharge_price=z["discharge_price"][i],
                setpoint=z["setpoint"][i], tr_infl=z["tr_infl"][i], tr_solar=z["tr_solar"][i],
                tr_max_power=z["tr_max_power"][i], tr_min_power=z["tr_min_power"][i],
                tr_load_fc=z["tr_load_fc"][i], tr_pv_fc=z["tr_pv_fc"][i],
                dr_start=z["dr_start"][i], dr_end=z["dr_end"][i], dr_cap=z["dr_cap"][i],
                dr_count=z["dr_count"][i], sessions=sess,
                grid_active=z["grid_active"][i] if "grid_active" in z else np.zeros((0, 0)),
                grid_reactive=z["grid_reactive"][i] if "grid_reactive" in z else np.zeros((0, 0)),
                date_feat=z["date_feat"][i] if "date_feat" in z else np.zeros((0, 3)),
                luts_c=z["luts_c"][l_off[i]:l_off[i + 1]], luts_d=z["luts_d"][l_off[i]:l_off[i + 1]],
            ).normalise())
        return cls(topo=topo, scenarios=out, config_name=str(z["config_name"]))


@dataclass
class SpawnTables:
    """What `EV_spawner` / `spawn_single_EV` (ev2gym/utilities/utils.py:477-557, 177-345) read besides their random
    draws: the arrival-rate, required-energy and time-of-stay tables of the config's scenario, the EV model table and a
    handful of config scalars.  Exported once from a reference env (reference_export.spawn_tables_from_env, stored next
    to the scenario banks by tools/make_golden.py --spawn-tables) so that the DEVICE sampler (ev2b_resample_sessions)
    needs neither the reference nor its data files."""
    workplace: int                  # 1: scenario == "workplace" (closed before 6 h / after 18 h and at weekends, :509-520)
    arrival_week: np.ndarray        # [96] df_arrival_week[scenario], per quarter of an hour (:515)
    arrival_weekend: np.ndarray     # [96] df_arrival_weekend[scenario] (:522)
    req_energy_mean: np.ndarray     # [48] df_req_energy[scenario] by half hour of arrival (:203-205)
    stay_mean: np.ndarray           # [48] df_time_of_stay_vs_arrival[scenario], hours (:231-234)
    spawn_multiplier: float         # config["spawn_multiplier"]
    min_stay_steps: int             # config["ev"]["min_time_of_stay"] // timescale (:495-496)
    desired_frac: float             # config["ev"]["desired_capacity"]
    min_battery_capacity: float
    min_emergency_battery_capacity: float
    ts_multiplier: float            # transition_soc_multiplier (1 when absent, :263-266)
    empty_ports_at_end: int         # empty_ports_at_end_of_simulation (:254-256)
    heterogeneous: int              # heterogeneous_ev_specs
    model_prob: np.ndarray          # [M] normalized_ev_registrations            (M = 1, prob 1 when homogeneous)
    model_B: np.ndarray             # [M] battery_capacity
    model_pmax_ac: np.ndarray       # [M] max_ac_charge_power
    model_pmax_dis: np.ndarray      # [M] max_discharge_power as the EV stores it (<= 0)
    model_pmin_ac: np.ndarray       # [M]
    model_pmin_dis: np.ndarray      # [M]
    model_phases: np.ndarray        # [M] int32
    model_lut: np.ndarray           # [M] int32 row of `luts` or -1 (scalar efficiencies drawn per EV, :290-296)
    luts: np.ndarray                # [L,101] percent; the reference uses the same curve for both directions (:288)
    power_setpoint_enabled: int = 0      # config["power_setpoint_enabled"]: setpoints are derived from the sessions (utils.py:664-757)
    power_setpoint_flexibility: float = 0.0   # config["power_setpoint_flexiblity"], percent
    homog_ts: float = 1.0           # homogeneous config: transition_soc, charge / discharge efficiency
    homog_eta_c: float = 1.0
    homog_eta_d: float = 1.0
    start: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), dtype=np.int32))   # [n,3] weekday, hour, minute
                                    # of sim_date at reset() for the scenarios of the bank this file goes with

    _SCALARS = ("workplace", "spawn_multiplier", "min_stay_steps", "desired_frac", "min_battery_capacity",
                "min_emergency_battery_capacity", "ts_multiplier", "empty_ports_at_end", "heterogeneous", "homog_ts",
                "homog_eta_c", "homog_eta_d", "power_setpoint_enabled", "power_setpoint_flexibility")
    _ARRAYS = ("arrival_week", "arrival_weekend", "req_energy_mean", "stay_mean", "model_prob", "model_B", "model_pmax_ac",
               "model_pmax_dis", "model_pmin_ac", "model_pmin_dis", "model_phases", "model_lut", "luts", "start")

    def save(self, path: str) -> None:
        np.savez_compressed(path, **{k: np.asarray(getattr(self, k)) for k in self._SCALARS + self._ARRAYS})

    @classmethod
    def load(cls, path: str) -> "SpawnTables":
        z = np.load(path, allow_pickle=False)
        kw = {k: z[k].item() for k in cls._SCALARS}
        kw.update({k: z[k] for k in cls._ARRAYS})
        return cls(**kw)
