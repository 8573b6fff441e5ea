#!/usr/bin/env python
"""A/B timing of step-kernel variants on one GPU, in one process (so a short GPU slot covers every variant):

  python tools/ab_kernels.py [--workloads c3,c2,c4] [--variants percharger,evl:G=1,evl:G=2:SORT=1] [--out gpurun_out/ab.json]

A variant is `kernel[:KEY=VAL]...`: kernel = percharger | evl (EV2B_KERNEL), every KEY=VAL is exported as EV2B_EVL_<KEY>
(or EV2B_<KEY> when KEY starts with a '!') before the handles are created -- the tuning knobs ev2b_create reads.
For every workload and variant it builds the same rotating env groups as bench.py (footprint > 2x L2) and reports
  * us per launch over WHOLE EPISODES (reset + T steps of every group, one CUDA graph per sweep) -- the bench window,
  * us per launch of the idle steps (< 1 EV per env connected) and of the busy steps, and of the busiest step,
from a one-round graph replayed step by step.  Also checks that the variants agree on the battery levels at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import L2_BYTES, WORKLOADS, algorithmic_bytes_per_env_step, load_pack   # noqa: E402


def set_variant(v):
    parts = v.split(":")
    kernel = {"evl": "evlist"}.get(parts[0], parts[0])
    for k in [k for k in os.environ if k.startswith("EV2B_EVL_")]:
        del os.environ[k]
    os.environ["EV2B_KERNEL"] = kernel
    for kv in parts[1:]:
        k, val = kv.split("=")
        os.environ[("EV2B_" + k[1:]) if k.startswith("!") else ("EV2B_EVL_" + k)] = val


def time_variant(torch, topo, pack, E, reward, state, variant, min_seconds, occ=None):
    from ev2gym_b200.engine import BatchedEngine
    set_variant(variant)
    dev = torch.device("cuda", 0)
    probe = BatchedEngine(topo, 1, reward=reward, state=state)
    D = probe.D
    probe.close()
    state_bytes = E * (28 * topo.P + 4 * topo.P + 4 * D + 150)
    NG = max(2, int(np.ceil(2.2 * L2_BYTES / state_bytes)))
    engines = []
    for g in range(NG):
        eng = BatchedEngine(topo, E, reward=reward, state=state)
        eng.load_scenarios(pack.scenarios)
        engines.append(eng)
    low = -1.0 if topo.v2g_enabled else 0.0
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    actions = [torch.rand((E, topo.P), device=dev, generator=gen) * (1.0 - low) + low for _ in range(NG)]
    T = topo.T

    def sweep():
        for g in range(NG):
            engines[g].reset()
        for r in range(T):
            for g in range(NG):
                engines[g].step(actions[(g + r) % NG])
    sweep()
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        sweep()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    ev0.record(); graph.replay(); ev1.record()
    torch.cuda.synchronize(dev)
    n = max(1, int(np.ceil(min_seconds * 1e3 / ev0.elapsed_time(ev1))))
    ev0.record()
    for _ in range(n):
        graph.replay()
    ev1.record()
    torch.cuda.synchronize(dev)
    us_episode = ev0.elapsed_time(ev1) * 1e3 / (n * NG * T)
    caps = engines[0].state_tensors()["port_cap"].cpu().numpy().copy()
    kpi = engines[0].state_tensors()["env_kpi"].cpu().numpy().copy()
    # per-step profile
    for g in range(NG):
        engines[g].reset()
    if occ is None:                                       # connected EVs per env before every step (first variant only)
        occ = []
        hot = engines[0].state_tensors()["port_hot"]
        for t in range(T):
            w0 = hot[..., 0]
            t_arr = ((w0 & 0xFFFF) ^ 0x8000) - 0x8000
            t_dep = (((w0 >> 16) & 0xFFFF) ^ 0x8000) - 0x8000
            occ.append(float(((t_arr <= t) & (t <= t_dep)).sum().item()) / E)
            engines[0].step(actions[t % NG])
        for g in range(NG):
            engines[g].reset()
    for g in range(NG):
        engines[g].step(actions[g])
    torch.cuda.synchronize(dev)
    rg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(rg):
        for g in range(NG):
            engines[g].step(actions[g])
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
    evs[0].record()
    for r in range(1, T):
        rg.replay()
        evs[r].record()
    torch.cuda.synchronize(dev)
    prof = [float("nan")] + [evs[r - 1].elapsed_time(evs[r]) * 1e3 / NG for r in range(1, T)]
    kl = engines[0].kernel_launches()
    for e in engines:
        e.close()
    return us_episode, n * NG * T, D, caps, kpi, kl, prof, occ


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c3,c2,c4")
    ap.add_argument("--min-seconds", type=float, default=0.25)
    ap.add_argument("--out", default="")
    ap.add_argument("--variants", default="percharger,evl:G=1,evl:G=2,evl:G=4")
    args = ap.parse_args()
    import torch
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    lines = []
    for wl in args.workloads.split(","):
        pack_name, E, reward, state, desc = WORKLOADS[wl]
        pack = load_pack(pack_name)
        topo = pack.topo
        ref, occ = None, None
        for v in args.variants.split(","):
            try:
                us, K, D, caps, kpi, kl, prof, occ = time_variant(torch, topo, pack, E, reward, state, v, args.min_seconds, occ)
            except Exception as exc:  # keep going: one variant failing must not lose the others' numbers
                line = {"workload": wl, "variant": v, "error": repr(exc)}
                lines.append(line)
                print(json.dumps(line), flush=True)
                continue
            b = algorithmic_bytes_per_env_step(topo, D)
            idle = [p for p, n in zip(prof, occ) if n < 1.0 and p == p]
            busy = [p for p, n in zip(prof, occ) if n >= 1.0 and p == p]
            line = {"workload": wl, "variant": v, "lib": os.path.basename(os.environ.get("EV2B_LIB", "libev2b.so")),
                    "us_per_launch_episode": round(us, 3), "launches": K, "envs": E,
                    "env_steps_per_s": E / (us * 1e-6), "roofline_frac": b * E / (us * 1e-6) / 1e9 / peak,
                    "idle_us": round(float(np.mean(idle)), 2) if idle else None, "idle_steps": len(idle),
                    "busy_us": round(float(np.mean(busy)), 2) if busy else None, "busy_steps": len(busy),
                    "busiest_us": round(float(np.nanmax(prof)), 2), "busiest_step": int(np.nanargmax(prof)),
                    "occupancy_mean": round(float(np.mean(occ)), 2), "kernel_launches": kl}
            if ref is None:
                ref = (caps, kpi)
            else:
                line["battery_levels_equal_first"] = bool(np.array_equal(caps, ref[0]))
                line["max_rel_kpi_diff"] = float(np.max(np.abs(kpi - ref[1]) / np.maximum(1e-9, np.abs(ref[1]))))
            lines.append(line)
            print(json.dumps(line), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
