#!/usr/bin/env python
"""A/B timing of the two step kernels on one GPU, in one process (so a short GPU slot covers every variant):

  python tools/ab_kernels.py [--workloads c3,c2,c4] [--steps 448] [--out gpurun_out/ab_kernels.json]

For every workload and every variant (step_kernel = "percharger"; evl_step_kernel with G = 1, 2, 4 warps per env) it
builds the same rotating env groups as bench.py (footprint > 2x L2), captures 4 rounds in a CUDA graph, replays it for
`--steps` launches and prints one JSON line: us per launch, env-steps/s, algorithmic GB/s and its fraction of the
measured HBM peak.  Also checks that the variants agree on the final battery levels of the timed episode prefix.
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


def time_variant(torch, topo, pack, E, reward, state, kernel, G, steps, opt=""):
    from ev2gym_b200.engine import BatchedEngine
    os.environ["EV2B_KERNEL"] = kernel
    os.environ["EV2B_EVL_STAGE"] = "1" if opt == "stage" else "0"          # opt-in experiments of ev2b_evlist.cuh
    os.environ["EV2B_EVL_PREFETCH"] = opt[2:] if opt.startswith("pf") else "0"
    if G:
        os.environ["EV2B_EVL_G"] = str(G)
    else:
        os.environ.pop("EV2B_EVL_G", None)
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
        eng.reset(scn_ids=[(g * E + e) % len(pack) for e in range(E)])
        engines.append(eng)
    low = -1.0 if topo.v2g_enabled else 0.0
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    actions = [torch.rand((E, topo.P), device=dev, generator=gen) * (1.0 - low) + low for _ in range(NG)]
    R = 4

    def round_(r):
        for g in range(NG):
            engines[g].step(actions[(g + r) % NG])
    for r in range(24):                      # into the busy part of the episode (ports fill up over the first steps)
        round_(r)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for r in range(R):
            round_(24 + r)
    n_rep = max(1, min(steps // (NG * R), (topo.T - 24 - R) // R))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    ev0.record()
    for _ in range(n_rep):
        graph.replay()
    ev1.record()
    torch.cuda.synchronize(dev)
    K = n_rep * NG * R
    us = ev0.elapsed_time(ev1) * 1e3 / K
    caps = engines[0].state_tensors()["port_cap"].cpu().numpy().copy()
    rew = engines[0].out["reward"].cpu().numpy().copy()
    kl = engines[0].kernel_launches()
    for e in engines:
        e.close()
    return us, K, D, caps, rew, kl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c3,c2,c4")
    ap.add_argument("--steps", type=int, default=448)
    ap.add_argument("--out", default="")
    ap.add_argument("--variants", default="percharger:0,evlist:1,evlist:2,evlist:4",
                    help="kernel:G[:opt],... (G = warps per env; opt = stage | pf1 | pf2 | pf3, see ev2b_evlist.cuh)")
    args = ap.parse_args()
    import torch
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    lines = []
    for wl in args.workloads.split(","):
        pack_name, E, reward, state, desc = WORKLOADS[wl]
        pack = load_pack(pack_name)
        topo = pack.topo
        ref = None
        for v in args.variants.split(","):
            kernel, G, opt = (v.split(":") + ["", ""])[:3]
            G = int(G or 0)
            try:
                us, K, D, caps, rew, kl = time_variant(torch, topo, pack, E, reward, state, kernel, G, args.steps, opt)
            except Exception as exc:  # keep going: one variant failing must not lose the others' numbers
                line = {"workload": wl, "kernel": kernel, "G": G, "opt": opt, "error": repr(exc)}
                lines.append(line)
                print(json.dumps(line), flush=True)
                continue
            b = algorithmic_bytes_per_env_step(topo, D)
            line = {"workload": wl, "kernel": kernel, "G": G, "opt": opt, "lib": os.path.basename(os.environ.get("EV2B_LIB", "libev2b.so")),
                    "us_per_launch": us, "launches": K, "envs": E,
                    "env_steps_per_s": E / (us * 1e-6), "algorithmic_GBps": b * E / (us * 1e-6) / 1e9,
                    "roofline_frac": b * E / (us * 1e-6) / 1e9 / peak, "kernel_launches": kl}
            if ref is None:
                ref = (caps, rew)
            else:
                line["battery_levels_equal_percharger"] = bool(np.array_equal(caps, ref[0]))
                line["max_rel_reward_diff"] = float(np.max(np.abs(rew - ref[1]) / np.maximum(1e-12, np.abs(ref[1]))))
            lines.append(line)
            print(json.dumps(line), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
