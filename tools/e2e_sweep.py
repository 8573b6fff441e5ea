#!/usr/bin/env python
"""e2e (ev2b_step_host, pinned host buffers) of the c3 workload for different numbers of pipelined env chunks
(EV2B_HOST_CHUNKS): whole episodes, us per step.  One JSON line per setting."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, load_pack   # noqa: E402


def main():
    import torch
    from ev2gym_b200.engine import BatchedEngine
    pack_name, E, reward, state, _ = WORKLOADS["c3"]
    pack = load_pack(pack_name)
    topo = pack.topo
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    for chunks in (1, 2, 3, 4, 6):
        os.environ["EV2B_HOST_CHUNKS"] = str(chunks)
        eng = BatchedEngine(topo, E, reward=reward, state=state)
        eng.load_scenarios(pack.scenarios)
        h_act = pin((E, topo.P), torch.float32)
        h_act[:] = np.random.default_rng(5).uniform(-1, 1.0, h_act.shape)
        h_rew, h_st, h_obs = pin((E,), torch.float64), pin((E,), torch.int32).view(np.uint32), pin((E, eng.D), torch.float32)
        res = {}
        for name, obs in (("with_obs", h_obs), ("reward_only", None)):
            for rep in range(3):
                eng.reset()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(topo.T):
                    eng.step_host(h_act, h_rew, h_st, obs)
                dt = time.perf_counter() - t0
            res[name + "_us_per_step"] = dt / topo.T * 1e6
        print(json.dumps({"chunks": chunks, **res, "env_steps_per_s_with_obs": E / (res["with_obs_us_per_step"] * 1e-6)}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
