"""Batched trajectory generator: the B200 counterpart of the reference's offline-RL data script
(ev2gym/scripts/generate_trajectories.py:14-105).

The reference builds trajectory i by `env.reset()`, picking ChargeAsFastAsPossible for even i and RoundRobin for odd i
("mixed-RR-Asap", :41,63-66), then looping `actions = agent.get_action(env); new_state, reward, done, _, _ =
env.step(actions)` and appending (state, actions, reward, done) (:69-81).  Two details of that loop are kept:
  * `observations[t]` is the state BEFORE step t (the reset observation first, :74,78);
  * `actions[t]` is appended AFTER `env.step` ran, and the step zeroes empty-port entries in the caller's array
    (ev_charger.py:137-140) -- so what is stored is the masked action.

Here E trajectories advance together: the agents' `get_action` is one kernel (ev2b_agent_actions), the step another,
and the [T,E,...] history stays on the device until the batch is finished.  Output format == the reference's pickle:
a list of dicts with "observations" [T,D], "actions" [T,P], "rewards" [T], "dones" [T].
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import BatchedEngine
from .scenario import Scenario, Topology

MIXED_RR_ASAP = ("afap", "roundrobin")      # trajectory i uses agents[i % len(agents)]  generate_trajectories.py:63-66


def generate_trajectories(topo: Topology, scenarios: Sequence[Scenario], n_trajectories: int,
                          agents: Sequence[str] = MIXED_RR_ASAP, reward: str = "SquaredTrackingErrorReward",
                          state: str = "PublicPST", device: int = 0, max_envs: int = 4096,
                          engine: Optional[BatchedEngine] = None) -> List[Dict[str, np.ndarray]]:
    """Roll `n_trajectories` full episodes; trajectory i plays scenario i mod len(scenarios) with agent
    agents[i mod len(agents)] ("afap", "roundrobin", "calap", "zero")."""
    if n_trajectories < 1:
        return []
    E = min(int(n_trajectories), int(max_envs))
    eng = engine or BatchedEngine(topo, E, reward=reward, state=state, device=device,
                                  outputs=("reward", "status", "obs", "action_mask"))
    if engine is None:
        eng.load_scenarios(list(scenarios))
    elif eng.E != E or not {"reward", "status", "obs", "action_mask"} <= set(eng.out):
        raise ValueError("engine must have n_envs == min(n_trajectories, max_envs) and outputs reward/status/obs/action_mask")
    torch = eng.torch
    T, P, D = topo.T, topo.P, eng.D
    kinds = list(agents)
    out: List[Dict[str, np.ndarray]] = []
    obs_h = torch.empty((T, E, D), dtype=torch.float32, device=eng.dev)
    act_h = torch.empty((T, E, P), dtype=torch.float64, device=eng.dev)
    rew_h = torch.empty((T, E), dtype=torch.float64, device=eng.dev)
    done_h = torch.empty((T, E), dtype=torch.bool, device=eng.dev)
    per_kind = {k: torch.empty((E, P), dtype=torch.float64, device=eng.dev) for k in kinds}
    for first in range(0, n_trajectories, E):
        idx = np.arange(first, first + E)                    # the tail batch rolls E envs too; extras are dropped below
        obs = eng.reset(scn_ids=idx % len(scenarios))
        which = torch.as_tensor(idx % len(kinds), device=eng.dev)
        occupied = torch.zeros((E, P), dtype=torch.float64, device=eng.dev)     # nothing is connected at t = 0
        for t in range(T):
            obs_h[t].copy_(obs)
            act = None
            for j, k in enumerate(kinds):
                a = eng.agent_actions(k, out=per_kind[k])
                act = a.clone() if act is None else torch.where((which == j)[:, None], a, act)
            res = eng.step(act)
            act_h[t] = act * occupied
            rew_h[t].copy_(res["reward"])
            done_h[t] = (res["status"] & 1) != 0
            occupied = res["action_mask"].to(torch.float64)
            obs = res["obs"]
        o, a, r, d = obs_h.cpu().numpy(), act_h.cpu().numpy(), rew_h.cpu().numpy(), done_h.cpu().numpy()
        for e in range(min(E, n_trajectories - first)):
            out.append({"observations": o[:, e].astype(np.float64), "actions": a[:, e].copy(), "rewards": r[:, e].copy(),
                        "dones": d[:, e].copy()})
    return out
