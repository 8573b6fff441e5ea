#!/usr/bin/env python
"""A tiny workload that touches every kernel of libev2b.so, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_run.py

event-driven kernel with 1 / 2 / 4 warps per env and one env per CTA, lean and HEAVY (statistics mode, per-port outputs,
histories, distribution grid), step_kernel, the k-step kernel with device-side reset, the agent kernel, the device
sampler's three kernels, reset kernels, episode statistics.  No checks beyond "it runs": parity is tests/'s job."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ev2gym_b200.engine import BatchedEngine
    from ev2gym_b200.scenario import SpawnTables, Topology
    from ev2gym_b200.synthetic import add_grid, sample_bank
    rng = np.random.default_rng(0)
    hist = ("hist_cs_power", "hist_cs_current", "hist_tr_overload", "hist_usage")
    heavy = ("dep_sat", "dep_cap", "port_energy")
    for kernel, G, tpb, n_ports, stats, outs in [("evlist", 1, 128, 2, False, hist), ("evlist", 2, 128, 2, True, heavy),
                                                 ("evlist", 4, 128, 1, False, ()), ("evlist", 1, 32, 1, True, heavy + hist),
                                                 ("percharger", 0, 128, 2, True, heavy + hist)]:
        os.environ["EV2B_KERNEL"], os.environ["EV2B_EVL_G"], os.environ["EV2B_EVL_TPB"] = kernel, str(G or 1), str(tpb)
        topo = Topology.uniform(C=12, n_ports=n_ports, Tr=3, T=16)
        bank = sample_bank(topo, 3, seed=1, min_stay=3)
        E = 6
        eng = BatchedEngine(topo, E, reward="ProfitMax_TrPenalty_UserIncentives", state="V2G_profit_max_loads", stats=stats,
                            outputs=("reward", "status", "obs", "action_mask", "cs_power", "tr_power", "tr_overload") + outs)
        eng.load_scenarios(bank)
        eng.reset()
        for t in range(topo.T):
            eng.step(torch.tensor(rng.uniform(-1, 1, (E, topo.P)), dtype=torch.float32, device="cuda"))
        if stats:
            eng.episode_stats()
        eng.reset_done()
        if not stats:
            eng.step_k(topo.T + 4, "uniform", seed=3, auto_reset=True)           # KSTEP kernel, device-side reset
        eng.step_k(3, "roundrobin", auto_reset=True)                             # agent kernel + launch per step
        eng.step_k(2, "calap")
        torch.cuda.synchronize()
        print("ok", kernel, G, tpb, n_ports, stats, eng.kernel_launches(), flush=True)
        eng.close()
    # distribution grid (HEAVY) on both kernels
    for kernel in ("evlist", "percharger"):
        os.environ["EV2B_KERNEL"], os.environ["EV2B_EVL_G"], os.environ["EV2B_EVL_TPB"] = kernel, "1", "128"
        topo = Topology.uniform(C=16, n_ports=1, Tr=6, T=12, imax=32.0)
        bank = sample_bank(topo, 2, seed=2, loads=False)
        add_grid(topo, bank, seed=2)
        eng = BatchedEngine(topo, 4, reward="V2G_grid_full_reward", state="V2G_grid_state", outputs=("reward", "status", "obs", "node_voltage"))
        eng.load_scenarios(bank)
        eng.reset()
        for t in range(topo.T):
            eng.step(torch.tensor(rng.uniform(-1, 1, (4, topo.P)), dtype=torch.float64, device="cuda"))
        torch.cuda.synchronize()
        print("ok grid", kernel, flush=True)
        eng.close()
    # device sampler
    from ev2gym_b200.scenario import ScenarioPack
    os.environ["EV2B_KERNEL"] = "evlist"
    pack = ScenarioPack.load(os.path.join(ROOT, "ev2gym_b200", "data", "c2_publicpst_c25.npz"))
    tab = SpawnTables.load(os.path.join(ROOT, "ev2gym_b200", "data", "spawn_c2_publicpst_c25.npz"))
    eng = BatchedEngine(pack.topo, 8, reward="SquaredTrackingErrorReward", state="PublicPST")
    eng.set_spawn_tables(tab)
    eng.load_scenarios(pack.scenarios[:8])
    eng.resample_sessions(seed=5)
    eng.reset()
    for t in range(30):
        eng.step(torch.rand((8, pack.topo.P), device="cuda"))
    torch.cuda.synchronize()
    print("ok sampler", len(eng.read_sessions(0)["port"]), "sessions in scenario 0", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
