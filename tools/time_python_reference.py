#!/usr/bin/env python
"""Times the UNMODIFIED Python reference (StavrosOrf/EV2Gym, installed into baseline/_ref by
`pip install --no-deps --target baseline/_ref /root/reference`, git-ignored) on this box's host cores:

    python tools/time_python_reference.py [--workload c3] [--seconds 8] [--procs N]

It steps `ev2gym.models.ev2gym_env.EV2Gym` -- the reference's own step() with its stock state and reward functions --
through whole episodes with uniform random actions, in one process and in N independent processes (the reference has no
vectorised env: N processes is how a user would use N cores), and prints ONE JSON line with env-steps/s (reset() time
excluded and reported separately).  gymnasium / matplotlib / pandapower are not installed: the stub packages of
oracle/refshim stand in for them (they are imported, never exercised by step()).  BASELINE.md section 3, items 1-2."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
CONFIGS = {  # workload -> (shipped config, size overrides, state fn, reward fn)   same shapes as bench.py's WORKLOADS
    "c2": ("PublicPST", {"number_of_charging_stations": 25}, "PublicPST", "SquaredTrackingErrorReward"),
    "c3": ("V2GProfitPlusLoads", {"number_of_charging_stations": 100, "number_of_ports_per_cs": 2, "number_of_transformers": 5},
           "V2G_profit_max_loads", "ProfitMax_TrPenalty_UserIncentives"),
    "c4": ("V2GProfitMax", {"number_of_charging_stations": 250}, "V2G_profit_max", "profit_maximization"),
}


def worker(args):
    workload, seconds, seed = args
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, REF)
    os.chdir(REF)                                   # shipped configs name ./ev2gym/data/... relative paths
    import numpy as np
    import yaml
    from ev2gym.models.ev2gym_env import EV2Gym
    from ev2gym.rl_agent import reward as ref_reward, state as ref_state
    base, ov, st, rw = CONFIGS[workload]
    cfg = yaml.safe_load(open(os.path.join(REF, "ev2gym", "example_config_files", base + ".yaml")))
    cfg.update(ov)
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    env = EV2Gym(config_file=f.name, seed=seed, state_function=getattr(ref_state, st), reward_function=getattr(ref_reward, rw))
    os.unlink(f.name)
    rng = np.random.default_rng(seed)
    low = -1.0 if env.config["v2g_enabled"] else 0.0
    steps, t_step, t_reset, episodes = 0, 0.0, 0.0, 0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end or episodes == 0:
        t0 = time.perf_counter()
        env.reset(seed=seed + episodes)
        t1 = time.perf_counter()
        for _ in range(env.simulation_length):
            env.step(rng.uniform(low, 1.0, env.number_of_ports))
        t2 = time.perf_counter()
        t_reset += t1 - t0; t_step += t2 - t1
        steps += env.simulation_length; episodes += 1
    return steps, t_step, t_reset, episodes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--procs", type=int, default=0, help="processes of the parallel leg (0 = all host threads)")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "ev2gym")):
        print(json.dumps({"unavailable": "baseline/_ref/ev2gym is not installed"}))
        return
    n = a.procs or os.cpu_count() or 1
    s1, ts1, tr1, ep1 = worker((a.workload, a.seconds, 100))
    with mp.get_context("spawn").Pool(n) as pool:
        t0 = time.perf_counter()
        res = pool.map(worker, [(a.workload, a.seconds, 1000 + 17 * i) for i in range(n)])
        wall = time.perf_counter() - t0
    print(json.dumps({
        "workload": a.workload, "what": "unmodified ev2gym.models.ev2gym_env.EV2Gym.step incl. stock state + reward functions, "
                                        "uniform actions, whole episodes; reset() excluded from the step rates",
        "one_process": {"env_steps_per_s": s1 / ts1, "episodes": ep1, "reset_s_per_episode": tr1 / ep1},
        "n_processes": {"procs": n, "env_steps_per_s": sum(r[0] / r[1] for r in res),
                        "env_steps_per_s_incl_reset": sum(r[0] for r in res) / max(max(r[1] + r[2] for r in res), 1e-9),
                        "episodes": sum(r[3] for r in res), "wall_s": wall},
    }))


if __name__ == "__main__":
    main()
