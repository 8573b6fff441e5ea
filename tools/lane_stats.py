#!/usr/bin/env python
"""Warp-lane statistics of the EV loop on a bench workload, from the oracle's state (no GPU): for every env-step, which of
the connected EVs charge / discharge / idle, which charging ones are saturated (pilot == max) and which are in the
constant-voltage stage (the `exp` branch) -- and how many warp-iterations of each model path the event-driven kernel
runs with the EVs in list order versus grouped by path.  Planning aid for DESIGN.md section 8.

    python tools/lane_stats.py [--workload c3] [--envs 64]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, load_pack   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--gt", type=int, default=64, help="threads per env group (32 * G)")
    args = ap.parse_args()
    from oracle.oracle import OracleBatch
    pack_name, _, reward, state, _ = WORKLOADS[args.workload]
    pack = load_pack(pack_name)
    topo = pack.topo
    E = min(args.envs, len(pack.scenarios))
    orc = OracleBatch(topo, pack.scenarios[:E], reward=reward, state=state)
    orc.reset()
    rng = np.random.default_rng(0)
    low = -1.0 if topo.v2g_enabled else 0.0
    tot = dict(steps=0, evs=0, iters=0, chg_iters=0, dis_iters=0, chg_sorted=0, dis_sorted=0, chg=0, dis=0, idle=0)
    busy = dict(tot)
    for t in range(topo.T):
        a = rng.uniform(low, 1.0, (E, topo.P))
        occ_before = orc.arr["port_session"] >= 0
        orc.step(a)
        e_now = orc.arr["port_cur_energy"]
        for e in range(E):
            ports = np.nonzero(occ_before[e])[0]
            n = len(ports)
            en = e_now[e][ports]
            chg, dis = en > 0, en < 0
            warps = [slice(i, i + 32) for i in range(0, n, 32)]
            rec = dict(steps=1, evs=n, iters=len(warps), chg=int(chg.sum()), dis=int(dis.sum()), idle=int(n - chg.sum() - dis.sum()),
                       chg_iters=sum(bool(chg[w].any()) for w in warps), dis_iters=sum(bool(dis[w].any()) for w in warps),
                       chg_sorted=-(-int(chg.sum()) // 32), dis_sorted=-(-int(dis.sum()) // 32))
            for k, v in rec.items():
                tot[k] += v
                if n > 0.3 * topo.P:
                    busy[k] += v
    for name, d in (("whole episode", tot), ("busy steps (> 30 % of the ports occupied)", busy)):
        s = max(1, d["steps"])
        print(f"{args.workload} {name}: env-steps {d['steps']}, EVs connected per env-step {d['evs'] / s:.1f}, "
              f"warp-iterations {d['iters'] / s:.2f}; charging {d['chg'] / s:.1f}, discharging {d['dis'] / s:.1f}, "
              f"no energy moved {d['idle'] / s:.1f}")
        print(f"    warp-iterations that run the charge path {d['chg_iters'] / s:.2f} (grouped by path: {d['chg_sorted'] / s:.2f}), "
              f"the discharge path {d['dis_iters'] / s:.2f} (grouped: {d['dis_sorted'] / s:.2f})")


if __name__ == "__main__":
    main()
