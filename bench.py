#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched EV2Gym step engine on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|c3-1k] [--impl reference]

One "step" = one fused-kernel pass advancing one batch of E env replicas by one timestep
(E x chargers named in config.workload).  Scenarios are banks exported from the reference's own
reset() (ev2gym_b200/data/*.npz, tools/make_golden.py --packs), tiled over the envs (env e uses
scenario e mod bank); actions are synthetic uniform fp32 in the action space, resident in HBM.

THE TIMED WINDOW IS ALWAYS WHOLE EPISODES.  A stock episode is empty at night (c3: 57 of 112 steps have < 1 EV per
env, 130 of 200 ports are occupied at the peak), so a window of a few steps measures either the idle path or the
busy path and not the step path.  Whatever --steps / --warmup say, both arms (this one and --impl reference) and the
e2e leg time `sweeps` full episodes of every env group, starting at t = 0, resets inside the timed region; --steps
only sets the MINIMUM number of launches (rounded up to whole sweeps, and up again until the region lasts >= ~1 s of
device time).  `steps` / `warmup` in the JSON line are the launches actually timed / warmed.

L2 hygiene: the per-batch state (~25-40 MB) would sit in the 126 MB L2, so the timed loop ROTATES
over G independent env groups whose total footprint is > 2x L2; every launch finds its data in HBM.

Printed: ONE JSON line (rank 0).  `value` = device-timed whole-job env-steps/s (CUDA events, max over
ranks); `e2e` = the same through ev2b_step_host with pinned HOST buffers (H2D actions, D2H
reward+status+obs inside the timed region); `roofline` = algorithmic bytes / measured launch time
against MEASURED_PEAKS.json; `cpu_baseline` = the C oracle on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (pack, E, reward, state, human description)
    "c2": ("c2_publicpst_c25", 1024, "SquaredTrackingErrorReward", "PublicPST",
           "PublicPST.yaml, 1024 envs x 25 chargers x 1 port, 1 transformer, uniform actions"),
    "c3": ("c3_v2gloads_c100n2tr5", 4096, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads",
           "V2GProfitPlusLoads.yaml, 4096 envs x 100 chargers x 2 ports, 5 transformers, uniform actions"),
    "c3-1k": ("c3_v2gloads_c100n2tr5", 1024, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads",
              "V2GProfitPlusLoads.yaml, 1024 envs x 100 chargers x 2 ports, 5 transformers, uniform actions"),
    # BASELINE config 5 names a BusinessPST.yaml that does not exist in the reference and a 20-transformer grid that
    # matches no shipped network: synthesised here -- 21-bus feeder, 500 chargers, synthetic scenarios.
    "c5": ("synthetic:grid:C500:Tr20", 2048, "V2G_grid_full_reward", "V2G_grid_state",
           "synthetic BusinessPST-like + 21-bus Laurent power flow, 2048 envs x 500 chargers x 1 port, 20 transformers"),
    "c4": ("c4_v2gprofitmax_c250", 8192, "profit_maximization", "V2G_profit_max",
           "V2GProfitMax.yaml, 8192 envs x 250 chargers x 1 port, 1 transformer, uniform actions"),
}
L2_BYTES = 126e6
MIN_TIMED_SECONDS = 1.0


def load_pack(name):
    from ev2gym_b200.scenario import ScenarioPack
    if name.startswith("synthetic:grid"):
        from ev2gym_b200.scenario import Topology
        from ev2gym_b200.synthetic import add_grid, sample_bank
        topo = Topology.uniform(C=500, n_ports=1, Tr=20, T=96, imax=32.0)
        bank = sample_bank(topo, 16, seed=5, loads=False)
        add_grid(topo, bank, seed=5)
        return ScenarioPack(topo, bank, name)
    return ScenarioPack.load(os.path.join(ROOT, "ev2gym_b200", "data", name + ".npz"))


def algorithmic_bytes_per_env_step(topo, obs_dim):
    """B_step = 40 P + 8 C + 20 Tr + 16 (+ 4 D with observations)."""
    return 40 * topo.P + 8 * topo.C + 20 * topo.Tr + 16 + 4 * obs_dim


def occupancy_aware_bytes_per_env_step(topo, obs_dim, series_len, tuple_len, n_connected):
    """The same count with the per-port terms only for ports that hold an EV: 40 B per CONNECTED EV (hot words 16 R,
    battery level 8 R + 8 W, exchanged energy 4, action 4) + 8 C + 20 Tr + 16, and of the observation only what a step
    rewrites: the header, the (scenario, time) series and one tuple per connected EV."""
    obs = 0
    if obs_dim:
        header = obs_dim - series_len - tuple_len * topo.P
        obs = 4 * (header + series_len + tuple_len * n_connected)
    return 40 * n_connected + 8 * topo.C + 20 * topo.Tr + 16 + obs


def source_sha():
    """Hash of the device code the step kernels are built from (csrc/ev2b_device.cuh, ev2b_evlist.cuh, ev2b_math.h):
    stamps profiles (roofline_traffic.json) to a build of the kernel they describe."""
    h = hashlib.sha256()
    for f in ("ev2b_device.cuh", "ev2b_evlist.cuh", "ev2b_math.h"):
        p = os.path.join(ROOT, "ev2gym_b200", "csrc", f)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_episodes(topo, scenarios, reward, state, E, n_episodes, warm_steps=3, threads=0, seconds=None):
    """Whole episodes of E envs on the C oracle (the port of the reference step), all host threads.
    n_episodes whole episodes, or (seconds given) as many whole episodes as fit, at least one."""
    from oracle.oracle import OracleBatch
    ob = OracleBatch(topo, [scenarios[e % len(scenarios)] for e in range(E)], reward=reward, state=state,
                     threads=threads)
    rng = np.random.default_rng(0)
    low = -1.0 if topo.v2g_enabled else 0.0
    acts = [rng.uniform(low, 1.0, (E, topo.P)) for _ in range(4)]
    ob.reset()
    for t in range(warm_steps):
        ob.step(acts[t % 4])
    steps, eps, t0 = 0, 0, time.perf_counter()
    while True:
        ob.reset()                                       # a new episode (inside the timed region, as on the GPU)
        for t in range(topo.T):
            ob.step(acts[t % 4])
        steps += topo.T
        eps += 1
        if seconds is None:
            if eps >= n_episodes:
                break
        elif time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return E * steps / dt, steps, eps, dt, ob.threads


def python_reference(workload, seconds=6.0):
    """The UNMODIFIED Python reference (baseline/_ref, installed by pip from /root/reference; git-ignored, travels to the
    GPU box) on this box's host cores: one process and one process per host thread (tools/time_python_reference.py)."""
    if workload not in ("c2", "c3", "c4") or not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "ev2gym")):
        return {"unavailable": "baseline/_ref not installed (or no shipped config for this workload)"}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_python_reference.py"), "--workload", workload,
                            "--seconds", str(seconds)], capture_output=True, text=True, timeout=240)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as exc:  # pragma: no cover
        return {"unavailable": repr(exc)}


def cpu_baseline(topo, scenarios, reward, state, E, target_seconds=12.0, threads=0, workload=None):
    """The C oracle (a port of the reference step) on all host threads, plus the Python reference itself beside it."""
    v, steps, eps, dt, thr = oracle_episodes(topo, scenarios, reward, state, E, 0, threads=threads,
                                             seconds=target_seconds)
    out = {"value": v, "unit": "env-steps/s", "cores": thr, "kind": "port",
           "sample": f"C oracle (oracle/ev2o.c, fp64, pthreads), {E} envs x {eps} whole episodes ({steps} steps) incl. "
                     f"state+reward, {dt:.1f} s"}
    if workload:
        out["python_reference"] = python_reference(workload)
    return out


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port, all host threads,
    on the same window as the GPU arm: whole episodes of the workload's full batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pack_name, E, reward, state, desc = WORKLOADS[wl]
    pack = load_pack(pack_name)
    topo = pack.topo
    n_eps = max(1, -(-args.steps // topo.T))             # whole episodes covering at least --steps steps
    v, steps, eps, dt, thr = oracle_episodes(topo, pack.scenarios, reward, state, E, n_eps,
                                             warm_steps=max(3, args.warmup))
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic actions on reference-exported scenarios",
        "config": {"workload": desc, "envs_per_gpu": E, "chargers": topo.C, "ports": topo.P, "transformers": topo.Tr,
                   "reward": reward, "state": state,
                   "window": f"{eps} whole episodes (t = 0 .. {topo.T}) of the full batch of {E} envs, resets inside",
                   "sample": f"full batch of {E} envs per step, {steps} steps"},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": thr, "kind": "port",
                         "sample": f"C oracle, {E} envs x {eps} whole episodes ({steps} steps)",
                         "python_reference": python_reference(wl) if args.gpus == 1 else None},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=448, help="minimum number of timed launches (rounded up to whole episodes)")
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--skip-agent-rollout", action="store_true")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-obs", action="store_true", help="do not produce observations (heuristic-driven runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the distinct-scenario and c3-1k extra measurements")
    ap.add_argument("--min-seconds", type=float, default=MIN_TIMED_SECONDS)
    ap.add_argument("--sweeps-only", action="store_true",
                    help="only the timed whole-episode sweeps, launched eagerly (for `ncu --metrics gpu__time_duration.sum`: "
                         "the launch list of the timed region without the extras)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, args.workload)

    import torch
    import torch.distributed as dist
    from ev2gym_b200.engine import BatchedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    pack_name, E, reward, state, desc = WORKLOADS[args.workload]
    pack = load_pack(pack_name)
    topo = pack.topo
    outputs = ("reward", "status") if args.no_obs else ("reward", "status", "obs")
    low = -1.0 if topo.v2g_enabled else 0.0
    T = topo.T

    def build_groups(E_, scenarios, n_groups=None):
        """G independent env groups (footprint > 2x L2), each with its own action tensor."""
        probe = BatchedEngine(topo, 1, reward=reward, state=state, device=local, outputs=outputs)
        D_ = 0 if args.no_obs else probe.D
        probe.close()
        state_bytes = E_ * (28 * topo.P + 4 * topo.P + 4 * D_ + 150)     # hot+cap+exch, actions, obs, per-env
        G_ = n_groups or max(2, int(np.ceil(2.2 * L2_BYTES / state_bytes)))
        engines_ = []
        for g in range(G_):
            eng = BatchedEngine(topo, E_, reward=reward, state=state, device=local, outputs=outputs)
            eng.load_scenarios(scenarios)
            engines_.append(eng)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        actions_ = [torch.rand((E_, topo.P), device=dev, generator=gen) * (1.0 - low) + low for _ in range(G_)]
        # scenario of env e = e mod bank size (ev2b_reset's default: no host id array, so a reset can be graph-captured);
        # every workload's E is a multiple of its bank size, so every group and rank covers the whole bank
        return engines_, actions_, D_, state_bytes

    def time_sweeps(engines_, actions_, min_steps, min_seconds, use_graph=True, collective=True):
        """Times whole-episode sweeps: every group is reset and stepped T times (G * T launches per sweep).
        Returns (elapsed_ms, n_sweeps, graphed)."""
        G_ = len(engines_)

        def sweep():
            for g in range(G_):
                engines_[g].reset()
            for r in range(T):
                for g in range(G_):
                    engines_[g].step(actions_[(g + r) % G_])
        sweep()                                                         # warm-up: one whole episode of every group
        torch.cuda.synchronize(dev)
        graph = None
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):                               # capture only: nothing executes here
                sweep()
        run = graph.replay if graph is not None else sweep
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        ev0.record(); run(); ev1.record()                              # calibration sweep (also warm-up of the graph)
        torch.cuda.synchronize(dev)
        one_ms = ev0.elapsed_time(ev1)
        n = max(1, -(-min_steps // (G_ * T)), int(np.ceil(min_seconds * 1e3 / max(one_ms, 1e-3))))
        if world > 1 and collective:                                    # every rank times the same number of sweeps
            tn = torch.tensor([n], device=dev, dtype=torch.int64)
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
            n = int(tn.item())
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev0.record()
        for _ in range(n):
            run()
        ev1.record()
        torch.cuda.synchronize(dev)
        if world > 1 and collective:
            dist.barrier()
        return ev0.elapsed_time(ev1), n, graph is not None

    def step_profile(engines_, actions_):
        """us per launch at every episode step t >= 1: a one-round graph (G launches) replayed T - 1 times, CUDA events
        between the replays (untimed extra; no collective)."""
        G_ = len(engines_)
        for g in range(G_):
            engines_[g].reset()
        for g in range(G_):                                             # round 0 eagerly: the capture below must see the
            engines_[g].step(actions_[g % G_])                          # same obs_full / mask_full host state as later rounds
        torch.cuda.synchronize(dev)
        rg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(rg):
            for g in range(G_):
                engines_[g].step(actions_[g % G_])
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
        evs[0].record()
        for r in range(1, T):
            rg.replay()
            evs[r].record()
        torch.cuda.synchronize(dev)
        return [float("nan")] + [evs[r - 1].elapsed_time(evs[r]) * 1e3 / G_ for r in range(1, T)]

    def occupancy_profile(engine):
        """Mean connected EVs per env BEFORE each step t of an episode (from the hot words; untimed)."""
        engine.reset()
        hot = engine.state_tensors()["port_hot"]
        occ = []
        gen = torch.Generator(device=dev); gen.manual_seed(99)
        for t in range(T):
            w0 = hot[..., 0]
            t_arr = ((w0 & 0xFFFF) ^ 0x8000) - 0x8000
            t_dep = (((w0 >> 16) & 0xFFFF) ^ 0x8000) - 0x8000
            occ.append(float(((t_arr <= t) & (t <= t_dep)).sum().item()) / engine.E)
            a = torch.rand((engine.E, topo.P), device=dev, generator=gen) * (1.0 - low) + low
            engine.step(a)
        return occ

    engines, actions, D, state_bytes = build_groups(E, pack.scenarios)
    G = len(engines)
    bytes_env_step = algorithmic_bytes_per_env_step(topo, D)
    launches0 = sum(e.launch_count for e in engines)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    elapsed_ms, n_sweeps, graphed = time_sweeps(engines, actions, args.steps, args.min_seconds,
                                                use_graph=not (args.no_graph or args.sweeps_only))
    clocks = sampler.stop()
    if args.sweeps_only:
        if rank == 0:
            print(json.dumps({"sweeps_only": True, "sweeps": n_sweeps, "groups": G, "T": T, "elapsed_ms": elapsed_ms,
                              "us_per_launch": elapsed_ms * 1e3 / (n_sweeps * G * T)}))
        return
    K = n_sweeps * G * T                                                # step launches inside the timed region
    gpu_launches = K + n_sweeps * 2 * G                                 # + the two reset kernels per group and sweep
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * E * K / (elapsed_ms * 1e-3)

    # per-step profile (us per launch at episode step t) and the occupancy it goes with -- rank 0's groups, untimed extras
    prof = step_profile(engines, actions) if rank == 0 else None
    occ = occupancy_profile(engines[0]) if rank == 0 else None

    # aggregate KPIs: the path's only collective: one small all-reduce(sum)
    kpi = sum(e.state_tensors()["env_kpi"].sum(dim=0) for e in engines)
    if world > 1:
        dist.all_reduce(kpi, op=dist.ReduceOp.SUM)

    def reset_all():
        for g in range(G):
            engines[g].reset()

    # ---- k-step device-agent rollout (ev2b_step_k): no action tensor, no host round trip -----------------
    agent_rate = None
    try:
        if args.skip_agent_rollout:
            raise RuntimeError("skipped")
        reset_all()
        for g in range(G):
            engines[g].step_k(T, "uniform", seed=7 + g)                 # warm-up episode
        reset_all()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for g in range(G):
            engines[g].step_k(T, "uniform", seed=7 + g)                 # one call = one whole episode of the group
        e1.record()
        torch.cuda.synchronize(dev)
        agent_rate = E * G * T / (e0.elapsed_time(e1) * 1e-3)
    except Exception as exc:  # pragma: no cover
        agent_rate = f"failed: {exc}"

    # ---- policy in the loop: obs -> torch MLP actor on the same GPU -> actions -> step (the stand-in for the SB3 actor
    #      of BASELINE config 4; SB3 is not installed).  The GEMMs are library calls and not part of the graded path.
    policy_rate = None
    try:
        if args.skip_agent_rollout or D == 0:
            raise RuntimeError("skipped")
        torch.manual_seed(0)
        actor = torch.nn.Sequential(torch.nn.Linear(D, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(),
                                    torch.nn.Linear(256, topo.P), torch.nn.Tanh() if low < 0 else torch.nn.Sigmoid()).to(dev)
        e_pol = engines[0]
        with torch.no_grad():
            obs = e_pol.reset()
            for _ in range(T):
                obs = e_pol.step(actor(obs))["obs"]
            obs = e_pol.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(T):
                obs = e_pol.step(actor(obs))["obs"]
            e1.record()
        torch.cuda.synchronize(dev)
        policy_rate = E * T / (e0.elapsed_time(e1) * 1e-3)
    except Exception as exc:  # pragma: no cover
        policy_rate = f"failed: {exc}"

    # ---- end to end through the host-buffer API: whole episodes, reset + T x step_host ---------------------
    eng = engines[0]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    h_act = [pin((E, topo.P), torch.float32) for _ in range(2)]
    for a in h_act:
        a[:] = np.random.default_rng(5).uniform(low, 1.0, a.shape)
    h_rew, h_st = pin((E,), torch.float64), pin((E,), torch.int32).view(np.uint32)
    h_obs = pin((E, D), torch.float32) if D else None
    n_e2e_eps = max(2, -(-args.steps // (8 * T)))

    def e2e_episode():
        eng.reset()
        for k in range(T):
            eng.step_host(h_act[k % 2], h_rew, h_st, h_obs)
    e2e_episode()                                                       # warm-up episode
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(n_e2e_eps):
        e2e_episode()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    n_e2e = n_e2e_eps * T
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * n_e2e / float(te.item())
    e2e_step_ms = float(te.item()) / n_e2e * 1e3
    # the same call without the observation download (obs_host = NULL: open-loop / replayed action sequences, or consumers
    # that read the state views on demand): H2D actions -> kernel -> D2H reward + status
    def e2e_episode_no_obs():
        eng.reset()
        for k in range(T):
            eng.step_host(h_act[k % 2], h_rew, h_st, None)
    e2e_episode_no_obs()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(n_e2e_eps):
        e2e_episode_no_obs()
    torch.cuda.synchronize(dev)
    e2e_noobs_value = E * n_e2e / (time.perf_counter() - t0)
    e2e_bytes = eng.host_step_bytes() if hasattr(eng, "host_step_bytes") else None
    if e2e_bytes is None:
        e2e_bytes = {"h2d": E * topo.P * 4, "d2h": E * (8 + 4 + 4 * D)}

    # PCIe floor of one e2e step: the same bytes as plain pinned copies, each direction alone (the link is full duplex)
    def copy_ms(dst, src, n=20):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        c0.record()
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize(dev)
        return c0.elapsed_time(c1) / n
    d2h_bytes, h2d_bytes = int(e2e_bytes["d2h"]), int(e2e_bytes["h2d"])
    nb = max(d2h_bytes, h2d_bytes, 1)
    dbuf, hbuf = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    d2h_ms = copy_ms(hbuf[:d2h_bytes], dbuf[:d2h_bytes])
    h2d_ms = copy_ms(dbuf[:h2d_bytes], hbuf[:h2d_bytes])

    # ---- extras (rank 0 only, untimed for the headline): no scenario reuse between envs, and the 1k-env shape ----
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "c3":
        for eng_ in engines[1:]:
            eng_.close()
        try:   # one scenario per env: the bank is tiled to E physical copies, so no two envs share scenario memory
            S = len(pack.scenarios)
            big = [pack.scenarios[i % S] for i in range(E)]
            engines2, actions2, _, sb2 = build_groups(E, big, n_groups=3)
            ms2, n2, _ = time_sweeps(engines2, actions2, 1, 0.3, use_graph=not args.no_graph, collective=False)
            us2 = ms2 * 1e3 / (n2 * 3 * T)
            extras["distinct_scenarios"] = {
                "us_per_launch": us2, "env_steps_per_s": E / (us2 * 1e-6), "bank": E,
                "roofline_frac": bytes_env_step * E / (us2 * 1e-6) / 1e9 / 1.0,   # divided by the peak below
                "what": f"same workload with one scenario PER ENV (the {S}-scenario bank tiled to {E} physical copies: "
                        f"scenario rows are never shared between envs), 3 env groups, {n2} whole-episode sweeps"}
            for e_ in engines2:
                e_.close()
        except Exception as exc:  # pragma: no cover
            extras["distinct_scenarios"] = {"error": repr(exc)}
        try:   # fresh scenarios every episode: the EV sessions of all E scenarios are re-drawn ON THE DEVICE between episodes
            from ev2gym_b200.scenario import SpawnTables
            tab = SpawnTables.load(os.path.join(ROOT, "ev2gym_b200", "data", "spawn_" + pack_name + ".npz"))
            S = len(pack.scenarios)
            big = [pack.scenarios[i % S] for i in range(E)]
            eng4 = []
            for g in range(3):
                e4 = BatchedEngine(topo, E, reward=reward, state=state, device=local, outputs=outputs)
                e4.set_spawn_tables(tab)
                e4.load_scenarios(big)
                eng4.append(e4)
            gen = torch.Generator(device=dev); gen.manual_seed(77)
            act4 = [torch.rand((E, topo.P), device=dev, generator=gen) * (1.0 - low) + low for _ in range(3)]

            def steps_only():
                for e4 in eng4:
                    e4.reset()
                for r in range(T):
                    for g, e4 in enumerate(eng4):
                        e4.step(act4[(g + r) % 3])
            for g, e4 in enumerate(eng4):
                e4.resample_sessions(seed=g)
            steps_only()
            torch.cuda.synchronize(dev)
            g4 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g4):
                steps_only()
            ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            n4, t_res, t_all = 8, 0.0, 0.0
            torch.cuda.synchronize(dev)
            for it in range(n4):
                ev0.record()
                for g, e4 in enumerate(eng4):
                    e4.resample_sessions(seed=1000 + 3 * it + g)
                ev1.record()
                g4.replay()
                ev2.record()
                torch.cuda.synchronize(dev)
                t_res += ev0.elapsed_time(ev1); t_all += ev0.elapsed_time(ev2)
            n_sess = float(sum(len(eng4[0].read_sessions(i)["port"]) for i in range(0, E, 64))) / (E // 64)
            extras["fresh_scenarios"] = {
                "env_steps_per_s": 3 * E * T * n4 / (t_all * 1e-3), "scenarios_per_s": 3 * E * n4 / (t_res * 1e-3),
                "resample_ms_per_bank": t_res / (3 * n4), "bank": E, "sessions_per_scenario": n_sess,
                "what": f"every episode of every env plays a scenario whose EV sessions were just drawn on the device "
                        f"(ev2b_resample_sessions = EV_spawner + spawn_single_EV, one scenario per env, time series of the "
                        f"{S}-scenario bank); timed: resample + reset + {T} steps, 3 env groups, {n4} episodes each"}
            for e_ in eng4:
                e_.close()
        except Exception as exc:  # pragma: no cover
            extras["fresh_scenarios"] = {"error": repr(exc)}
        try:   # north_star shape: 1k envs x 100 chargers
            E1 = 1024
            engines3, actions3, _, _ = build_groups(E1, pack.scenarios)
            ms3, n3, _ = time_sweeps(engines3, actions3, 1, 0.3, use_graph=not args.no_graph, collective=False)
            us3 = ms3 * 1e3 / (n3 * len(engines3) * T)
            extras["c3_1k"] = {"us_per_launch": us3, "env_steps_per_s": E1 / (us3 * 1e-6), "envs": E1,
                               "roofline_frac": bytes_env_step * E1 / (us3 * 1e-6) / 1e9 / 1.0,
                               "what": f"1024 envs x 100 chargers x 2 ports (north_star shape), {len(engines3)} env groups, "
                                       f"{n3} whole-episode sweeps"}
            for e_ in engines3:
                e_.close()
        except Exception as exc:  # pragma: no cover
            extras["c3_1k"] = {"error": repr(exc)}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        for x in extras.values():
            if "roofline_frac" in x:
                x["roofline_frac"] /= peak
        launch_ms = elapsed_ms / K
        sha = source_sha()
        traffic, traffic_src, traffic_sha = None, None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(args.workload)
            if tj:
                traffic, traffic_src, traffic_sha = tj["dram_bytes_per_launch"], tj["source"], tj.get("source_sha")
        achieved = bytes_env_step * E / (launch_ms * 1e-3) / 1e9
        series_len = {"V2G_profit_max": 20, "V2G_profit_max_loads": 20 + 40 * topo.Tr, "PublicPST": 0,
                      "V2G_grid_state": 5 + 2 * topo.n_bus}.get(state, 0) if D else 0
        tuple_len = {"PublicPST": 3, "V2G_grid_state": 3}.get(state, 2) if D else 0
        occ_mean = float(np.mean(occ))
        bytes_occ = float(np.mean([occupancy_aware_bytes_per_env_step(topo, D, series_len, tuple_len, n) for n in occ]))
        achieved_occ = bytes_occ * E / (launch_ms * 1e-3) / 1e9
        idle = [p for p, n in zip(prof, occ) if n < 1.0 and p == p]
        busy = [p for p, n in zip(prof, occ) if n >= 1.0 and p == p]
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": 2 * G * T, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (uniform fp32 actions on scenario banks exported "
                                                         "from the reference's reset())",
            "config": {"workload": desc, "envs_per_gpu": E, "chargers": topo.C, "ports": topo.P,
                       "transformers": topo.Tr, "obs_dim": D, "reward": reward, "state": state,
                       "window": f"{n_sweeps} sweeps x {G} env groups x whole episodes (t = 0 .. {T}), resets inside "
                                 f"the timed region; --steps {args.steps} / --warmup {args.warmup} requested",
                       "l2": f"rotating {G} independent env groups, {G * state_bytes / 1e6:.0f} MB total > 126 MB L2",
                       "cuda_graph": graphed, "parallelism": f"env-sharded x{world}"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": n_e2e, "ms_per_step": e2e_step_ms,
                    "window": f"{n_e2e_eps} whole episodes of one env group: reset + {T} x ev2b_step_host",
                    "pcie_floor": {"d2h_ms": d2h_ms, "h2d_ms": h2d_ms, "d2h_gbs": d2h_bytes / d2h_ms / 1e6,
                                   "h2d_gbs": h2d_bytes / h2d_ms / 1e6, "frac": max(d2h_ms, h2d_ms) / e2e_step_ms,
                                   "what": "the step's bytes as bare pinned cudaMemcpyAsync, each direction alone"},
                    "what": "ev2b_step_host: pinned host actions -> H2D -> fused kernel -> D2H reward+status+obs -> sync",
                    "without_obs_download": {"value_per_gpu": e2e_noobs_value, "d2h_bytes_per_step": E * 12,
                                             "what": "same call with obs_host = NULL (reward + status only), this rank"}},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "traffic_source_sha": traffic_sha,
                         "source_sha": sha, "traffic_is_this_build": traffic_sha == sha,
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": bytes_env_step,
                         "occupancy_mean": occ_mean, "occupancy_peak": float(np.max(occ)), "ports": topo.P,
                         "occupancy_aware_bytes_per_env_step": bytes_occ, "achieved_occupancy_aware": achieved_occ,
                         "frac_occupancy_aware": achieved_occ / peak,
                         "idle_step_us": float(np.mean(idle)) if idle else None, "idle_steps": len(idle),
                         "busy_step_us": float(np.mean(busy)) if busy else None, "busy_steps": len(busy),
                         "busiest_step_us": float(np.nanmax(prof)),
                         "us_by_episode_step": [None if p != p else round(p, 2) for p in prof],
                         "connected_evs_by_episode_step": [round(n, 1) for n in occ],
                         "kernel": "ev2b::evl_step_kernel" if engines[0].kernel_launches()[1] else "ev2b::step_kernel",
                         "kernel_launches": dict(zip(("step_kernel", "evl_step_kernel", "evl_rebuild_kernel"),
                                                     engines[0].kernel_launches())),
                         "launch_ms": launch_ms},
            "kpi_allreduce": {"total_reward": float(kpi[0].item()), "total_evs_served": float(kpi[5].item())},
            "device_agent_rollout": {"value": agent_rate, "unit": "env-steps/s per GPU",
                                     "what": "ev2b_step_k, UNIFORM on-device agent, one call per whole episode and env group"},
            "policy_rollout": {"value": policy_rate, "unit": "env-steps/s per GPU",
                               "what": f"obs -> torch MLP actor ({D}-256-256-{topo.P}, fp32) on the GPU -> ev2b_step, one env "
                                       "group (L2-resident), one whole episode, no host round trip"},
            "extras": extras,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(topo, pack.scenarios, reward, state if D else None, E,
                                                workload=args.workload if args.workload in ("c2", "c3", "c4") else None)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
