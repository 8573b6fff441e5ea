"""Synthetic scenario sampler (numpy).  Produces `Scenario`s with the same structure and value
ranges as the reference's `reset()` (EV_spawner / spawn_single_EV, ev2gym/utilities/utils.py:177-345,
477-557) without the reference's data files: arrival probabilities, stay lengths and EV models are
drawn from simple parametric distributions.  Used for tests at arbitrary sizes and as a fallback
scenario source when no reference-exported scenario pack is available.  Not a port of the reference
generators: scenarios exported from the reference itself (tools/make_golden.py --packs) are what the
benchmark and the parity fixtures use.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .scenario import LUT_LEN, Scenario, Topology

# (battery kWh, max AC kW, max discharge kW, has efficiency table)
_MODELS = [(57.5, 11.0, 11.0, True), (77.0, 11.0, 11.0, True), (64.0, 7.4, 7.4, True), (40.0, 6.6, 6.6, False),
           (82.0, 22.0, 0.0, False), (28.5, 3.7, 0.0, False), (50.0, 11.0, 11.0, False)]


def _lut(rng) -> np.ndarray:
    base = rng.uniform(75, 92)
    t = np.clip(base + np.cumsum(rng.uniform(-0.5, 1.5, LUT_LEN)) * 0.2, 60, 99)
    return np.round(t, 1)


def sample_scenario(topo: Topology, rng: np.random.Generator, occupancy: float = 0.5,
                    heterogeneous: bool = True, setpoints: bool = True, loads: bool = True,
                    min_stay: int = 8, two_stage: bool = True) -> Scenario:
    T, Tr, C, P = topo.T, topo.Tr, topo.C, topo.P
    port_cs = np.repeat(np.arange(C), topo.cs_n_ports)
    n_lut = 3 if heterogeneous else 0
    luts = np.stack([_lut(rng) for _ in range(n_lut)]) if n_lut else np.ones((0, LUT_LEN))
    rows = []
    for p in range(P):                 # spawner-port timeline; gaps >= 3 steps (utils.py:534-536)
        t = 2 + int(rng.integers(0, max(2, int(8 / max(occupancy, 0.05)))))
        while True:
            stay = int(max(min_stay, rng.normal(24, 10)))
            t_arr, t_dep = t + 1, t + stay + 3
            if t_dep + 1 >= T:         # empty_ports_at_end_of_simulation (utils.py:254-256)
                break
            rows.append((t_arr, t_dep, port_cs[p], p))
            t = t_dep + 2 + int(rng.geometric(min(0.9, max(0.02, occupancy / 6))))
    rows.sort(key=lambda r: (r[0], r[3]))
    S = len(rows)
    sess = {k: np.zeros(S) for k in ("cap0", "B", "pmax_ac", "pmin_ac", "pmax_dis", "pmin_dis", "bmin", "bmin_em",
                                     "desired", "ts", "mult", "eta_c", "eta_d")}
    sess.update({k: np.zeros(S, dtype=np.int32) for k in ("loc", "t_arr", "t_dep", "ev_phases", "lut")})
    for i, (ta, td, c, _) in enumerate(rows):
        m = _MODELS[int(rng.integers(len(_MODELS)))] if heterogeneous else (50.0, 11.0, 11.0, False)
        B = m[0]
        sess["loc"][i], sess["t_arr"][i], sess["t_dep"][i] = c, ta, td
        sess["B"][i], sess["pmax_ac"][i], sess["pmax_dis"][i] = B, m[1], -m[2]
        sess["pmin_ac"][i] = 0.0 if rng.random() < 0.8 else 1.4
        sess["pmin_dis"][i] = 0.0
        sess["bmin"][i] = 5.0
        sess["bmin_em"][i] = 25.0 if B >= 25 else 0.7 * B
        sess["desired"][i] = (1.0 if rng.random() < 0.7 else 0.8) * B
        req = max(5.0, rng.normal(0.45 * B, 0.2 * B))
        cap0 = B - req if req < B else float(rng.integers(1, int(B)))
        sess["cap0"][i] = max(cap0, 5.0) if B > 10 else cap0
        sess["ev_phases"][i] = 3 if rng.random() < 0.85 else 1
        sess["mult"][i] = 5.0
        if heterogeneous:
            sess["ts"][i] = np.round(0.9 - (rng.random() + 0.00001) / 5, 3) if two_stage else 1.0
            if m[3] and n_lut:
                sess["lut"][i] = int(rng.integers(n_lut))
                sess["eta_c"][i] = sess["eta_d"][i] = np.nan
            else:
                sess["lut"][i] = -1
                sess["eta_c"][i] = np.round(1 - (rng.random() + 0.00001) / 20, 3)
                sess["eta_d"][i] = np.round(1 - (rng.random() + 0.00001) / 20, 3)
        else:
            sess["ts"][i], sess["lut"][i], sess["eta_c"][i], sess["eta_d"][i] = 1.0, -1, 1.0, 1.0
    price = np.abs(rng.normal(0.12, 0.05, T // 4 + 1)).repeat(4)[:T] + 0.01
    max_power = np.full((Tr, T), 60.0 * max(1, C // max(Tr, 1)) * 0.35)
    dr_start = np.zeros((Tr, 1), dtype=np.int32)
    dr_end = np.zeros((Tr, 1), dtype=np.int32)
    dr_cap = np.zeros((Tr, 1))
    dr_count = np.zeros(Tr, dtype=np.int32)
    infl = np.zeros((Tr, T))
    solar = np.zeros((Tr, T))
    lfc = np.zeros((Tr, T))
    pfc = np.zeros((Tr, T))
    if loads:
        x = np.linspace(0, 2 * np.pi, T)
        for k in range(Tr):
            infl[k] = max_power[k] * np.clip(0.35 + 0.25 * np.sin(x + rng.uniform(0, 6)) + rng.normal(0, 0.05, T), 0, 1)
            solar[k] = -max_power[k] * np.clip(0.5 * np.sin(x * 0.5) ** 2 * rng.uniform(0.5, 1.1), 0, 1)
            lfc[k] = np.clip(rng.normal(0.3 * infl[k], 0.05 * np.abs(infl[k]) + 1e-9), -max_power[k], max_power[k])
            pfc[k] = rng.normal(0.2 * solar[k], 0.05 * np.abs(solar[k]) + 1e-9)
            s0 = int(rng.integers(5, max(6, T - 10)))
            dr_start[k, 0], dr_end[k, 0], dr_cap[k, 0], dr_count[k] = s0, s0 + 4, float(np.clip(rng.normal(35, 5), 0, 100)), 1
            max_power[k, s0:s0 + 4] *= 1 - dr_cap[k, 0] / 100
    setpoint = np.zeros(T)
    if setpoints:
        setpoint = np.clip(rng.normal(0.25, 0.15, T), 0, None) * P * 3.0
    return Scenario(charge_price=-price, discharge_price=price * 1.0, setpoint=setpoint,
                    tr_infl=infl, tr_solar=solar, tr_max_power=max_power, tr_min_power=-np.abs(max_power).max() *
                    np.ones((Tr, T)), tr_load_fc=lfc, tr_pv_fc=pfc, dr_start=dr_start, dr_end=dr_end, dr_cap=dr_cap,
                    dr_count=dr_count, sessions=sess, luts_c=luts, luts_d=luts.copy()).normalise()


def add_grid(topo: Topology, scenarios: List[Scenario], seed: int = 0) -> Topology:
    """Attach a synthetic radial feeder (one bus per transformer) to `topo` and base bus powers / calendar features
    to every scenario: a diagonally dominant K and L = 1 give a well-conditioned Laurent iteration."""
    rng = np.random.default_rng(seed)
    n, T = topo.Tr, topo.T
    path = np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
    K = -(0.02 + 0.01 * rng.random((n, n))) * np.exp(-0.15 * path) * (1 + 0.6j)
    topo.grid_K, topo.grid_L, topo.grid_s_base = K.astype(np.complex128), np.ones(n, dtype=np.complex128), 1000.0
    for sc in scenarios:
        x = np.linspace(0, 2 * np.pi, T + 1)[:, None]
        sc.grid_active = np.round(120 + 80 * np.sin(x + rng.uniform(0, 6, (1, n))) + 20 * rng.random((T + 1, n)), 1)
        sc.grid_reactive = np.round(0.4 * sc.grid_active, 1)
        hours = (5 + np.arange(T + 1) * topo.timescale // 60) % 24
        sc.date_feat = np.stack([np.full(T + 1, 2 / 7), np.sin(hours / 24 * 2 * np.pi), np.cos(hours / 24 * 2 * np.pi)], 1)
        sc.normalise()
    return topo


def sample_bank(topo: Topology, n: int, seed: int = 0, **kw) -> List[Scenario]:
    rng = np.random.default_rng(seed)
    return [sample_scenario(topo, rng, **kw) for _ in range(n)]
