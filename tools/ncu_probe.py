#!/usr/bin/env python
"""Workload for Nsight Compute: steps ONE c3-sized env group through its first `--steps` steps with each step-kernel
variant in turn, so a single ncu run sees step_kernel and evl_step_kernel (G = 1, 2, 4) on the same inputs:

  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,... --clock-control none --csv \
      --kernel-name regex:step_kernel --log-file gpurun_out/probe.csv python tools/ncu_probe.py --steps 32

(launch index within a variant == episode step; steps >= ~25 are the busy part of the episode).  Numbers printed by a
run under ncu are not bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, load_pack   # noqa: E402
from tools.ab_kernels import set_variant   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--variants", default="percharger,evl:G=1,evl:G=2,evl:G=4", help="tools/ab_kernels.py syntax")
    args = ap.parse_args()
    import torch
    from ev2gym_b200.engine import BatchedEngine
    pack_name, E, reward, state, desc = WORKLOADS[args.workload]
    pack = load_pack(pack_name)
    topo = pack.topo
    dev = torch.device("cuda", 0)
    low = -1.0 if topo.v2g_enabled else 0.0
    for v in args.variants.split(","):
        set_variant(v)
        eng = BatchedEngine(topo, E, reward=reward, state=state)
        eng.load_scenarios(pack.scenarios)
        eng.reset()
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234)
        for t in range(args.steps):
            a = torch.rand((E, topo.P), device=dev, generator=gen) * (1.0 - low) + low
            eng.step(a)
        torch.cuda.synchronize(dev)
        print(v, "launches", eng.kernel_launches(), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
