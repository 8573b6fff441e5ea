"""Attribute-compatible, read-only views of the engine state for ONE env.

The reference's plugin functions (ev2gym/rl_agent/state.py, reward.py, cost.py), heuristics
(ev2gym/baselines/heuristics.py) and wrappers never call an API: they reach into
`env.charging_stations[i].evs_connected[j].<attr>` and `env.transformers[k].<attr>`
(attribute contract: SURVEY.md section 8b).  These classes present the struct-of-arrays state a
step of the CUDA engine left behind under exactly those names, so such code runs unchanged on
`EV2GymB200`.  They are rebuilt from a host snapshot after every step; nothing here computes the
simulation itself.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np


class EVView:
    """Looks like ev2gym.models.ev.EV (attributes of ev.py:68-113) for a connected session."""

    def __init__(self, sc, i: int, port_in_cs: int, cap: float, exch: float, energy: float, amps: float,
                 timescale: int):
        s = sc.sessions
        self._i = i
        self.id = port_in_cs
        self.location = int(s["loc"][i])
        self.timescale = timescale
        self.time_of_arrival = int(s["t_arr"][i])
        self.time_of_departure = int(s["t_dep"][i])
        self.desired_capacity = float(s["desired"][i])
        self.battery_capacity_at_arrival = float(s["cap0"][i])
        self.battery_capacity = float(s["B"][i])
        self.min_battery_capacity = float(s["bmin"][i])
        self.min_emergency_battery_capacity = float(s["bmin_em"][i])
        self.max_ac_charge_power = float(s["pmax_ac"][i])
        self.min_ac_charge_power = float(s["pmin_ac"][i])
        self.max_discharge_power = float(s["pmax_dis"][i])
        self.min_discharge_power = float(s["pmin_dis"][i])
        self.transition_soc = float(s["ts"][i])
        self.transition_soc_multiplier = float(s["mult"][i])
        self.ev_phases = int(s["ev_phases"][i])
        lut = int(s["lut"][i])
        if lut >= 0:
            self.charge_efficiency = {k: float(v) for k, v in enumerate(sc.luts_c[lut])}
            self.discharge_efficiency = {k: float(v) for k, v in enumerate(sc.luts_d[lut])}
        else:
            self.charge_efficiency = float(s["eta_c"][i])
            self.discharge_efficiency = float(s["eta_d"][i])
        self.current_capacity = float(cap)
        self.total_energy_exchanged = float(exch)
        self.current_energy = float(energy)
        self.actual_current = float(amps)
        self.required_energy = self.battery_capacity - self.current_capacity

    def get_soc(self) -> float:                      # ev.py:223-229
        return self.current_capacity / self.battery_capacity

    def get_user_satisfaction(self) -> float:        # ev.py:204-214
        if self.current_capacity < self.desired_capacity - 0.001:
            return self.current_capacity / self.desired_capacity
        return 1

    def is_departing(self, timestep):                # ev.py:191-202
        return None if timestep < self.time_of_departure else self.get_user_satisfaction()


class ChargerView:
    """Looks like ev2gym.models.ev_charger.EV_Charger (ev_charger.py:41-94)."""

    def __init__(self, topo, c: int):
        self.id = c
        self.connected_bus = int(topo.cs_tr[c])
        self.connected_transformer = int(topo.cs_tr[c])
        self.n_ports = int(topo.cs_n_ports[c])
        self.charger_type = "AC"
        self.timescale = topo.timescale
        self.min_charge_current = float(topo.cs_imin[c])
        self.max_charge_current = float(topo.cs_imax[c])
        self.min_discharge_current = float(topo.cs_imin_dis[c])
        self.max_discharge_current = float(topo.cs_imax_dis[c])
        self.phases = int(topo.cs_phases[c])
        self.voltage = float(topo.cs_voltage[c])
        self.evs_connected: List[Optional[EVView]] = [None] * self.n_ports
        self.n_evs_connected = 0
        self.current_power_output = 0.0
        self.current_total_amps = 0.0
        self.current_signal = [0] * self.n_ports
        self.current_step = 0
        self.current_charge_price = 0.0
        self.current_discharge_price = 0.0
        self.total_energy_charged = 0.0
        self.total_energy_discharged = 0.0
        self.total_profits = 0.0
        self.total_evs_served = 0
        self.total_user_satisfaction = 0.0
        self.all_user_satisfaction: List[float] = []

    def get_max_power(self):                         # ev_charger.py:251-252
        return self.max_charge_current * self.voltage * math.sqrt(self.phases) / 1000

    def get_min_charge_power(self):                  # ev_charger.py:254-255
        return self.min_charge_current * self.voltage * math.sqrt(self.phases) / 1000

    def get_min_power(self):                         # ev_charger.py:257-258
        return self.max_discharge_current * self.voltage * math.sqrt(self.phases) / 1000

    def get_avg_user_satisfaction(self):             # ev_charger.py:260-264
        return 0 if self.total_evs_served == 0 else self.total_user_satisfaction / self.total_evs_served


class TransformerView:
    """Looks like ev2gym.models.transformer.Transformer (transformer.py:15-78, 142-188, 258-302)."""

    def __init__(self, topo, sc, k: int):
        self.id = k
        self.voltage = float(topo.tr_voltage)
        self.max_power = sc.tr_max_power[k]
        self.min_power = sc.tr_min_power[k]
        self.max_current = self.max_power * 1000 / self.voltage
        self.min_current = self.min_power * 1000 / self.voltage
        self.inflexible_load = sc.tr_infl[k]
        self.solar_power = sc.tr_solar[k]
        self.inflexible_load_forecast = sc.tr_load_fc[k].copy()
        self.pv_generation_forecast = sc.tr_pv_fc[k].copy()
        self.cs_ids = np.nonzero(topo.cs_tr == k)[0]
        self.simulation_length = topo.T
        self.steps_ahead = topo.dr_steps_ahead
        self.dr_events = [{"event_start_step": int(sc.dr_start[k, j]), "event_end_step": int(sc.dr_end[k, j]),
                           "capacity_percentage": float(sc.dr_cap[k, j])} for j in range(int(sc.dr_count[k]))]
        self.current_step = 0
        self.current_power = float(self.inflexible_load[0] + self.solar_power[0])
        self.current_amps = self.current_power * 1000 / self.voltage

    def is_overloaded(self) -> bool:                 # transformer.py:276-290
        e = 0.0001
        return bool(self.current_power > self.max_power[self.current_step] + e or
                    self.current_power < self.min_power[self.current_step] - e)

    def get_how_overloaded(self) -> float:           # transformer.py:292-302
        return float(np.abs(self.current_power - self.max_power[self.current_step])) if self.is_overloaded() else 0

    def get_power_limits(self, step, horizon):       # transformer.py:142-171
        known = max(self.max_power) * np.ones(horizon)
        limit = max(self.max_power)
        for ev in self.dr_events:
            if step + self.steps_ahead >= ev["event_start_step"] and ev["event_end_step"] >= step:
                v = limit - limit * ev["capacity_percentage"] / 100
                if step > ev["event_start_step"]:
                    known[:ev["event_end_step"] - step] = v
                else:
                    known[abs(ev["event_start_step"] - step):abs(ev["event_end_step"] - step)] = v
        return known

    def get_load_pv_forecast(self, step, horizon):   # transformer.py:173-188 (incl. the write-through)
        load = self.inflexible_load_forecast[step:step + horizon]
        pv = self.pv_generation_forecast[step:step + horizon]
        if step < len(self.inflexible_load_forecast):
            load[0] = self.inflexible_load[step]
            pv[0] = self.solar_power[step]
        if len(load) < horizon:
            load = np.append(load, np.ones(horizon - len(load)) * self.inflexible_load_forecast[-1])
            pv = np.append(pv, np.ones(horizon - len(pv)) * self.pv_generation_forecast[-1])
        return load, pv
