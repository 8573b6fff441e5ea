"""CPU restatement of the reference's stock heuristic agents -- TEST INFRASTRUCTURE ONLY (see oracle/README in
DESIGN.md section 3): used by tests/ to check the device agents of ev2b_agent_actions / ev2b_step_k.

Each class follows /root/reference/ev2gym/baselines/heuristics.py on an `OracleEnv` (oracle/oracle.py) instead of the
reference's object graph; list semantics (insert-at-front, remove, rotate) are kept literal so that the queue order --
and therefore which EVs are served each round -- is the reference's.  Pinned against the reference's own agents by the
golden traces tests/golden/*roundrobin*, *calap* (tools/make_golden.py records `agent.get_action(env)` per step).
"""
from __future__ import annotations

import math

import numpy as np


def _ports(env):
    """(charger id, local port) of every flat port index, in the reference's iteration order."""
    n = env.topo.cs_n_ports
    return [(c, j) for c in range(env.topo.C) for j in range(int(n[c]))]


def _soc(env, port: int) -> float:
    """EV.get_soc (ev.py:223-229) of the EV on `port`."""
    s = int(env.arr["port_session"][port])
    return float(env.arr["port_cap"][port]) / float(env.scenario.sessions["B"][s])


class OracleRoundRobin:
    """heuristics.py:7-95.  `average_power` in W per port (:19-23); the queue holds flat port indices."""

    def __init__(self, env):
        tp = env.topo
        acc = 0.0
        for c in range(tp.C):                                                        # :20-23
            acc += float(tp.cs_imax[c]) * float(tp.cs_voltage[c]) * math.sqrt(int(tp.cs_phases[c])) / int(tp.cs_n_ports[c])
        self.average_power = acc / tp.C
        self.ports_per_cs = int(tp.cs_n_ports[0])                                    # env.number_of_ports_per_cs
        self.queue = []

    def _refresh(self, env):                                                         # update_ev_buffer :33-52
        for port in range(env.topo.P):
            waiting = env.arr["port_session"][port] >= 0 and _soc(env, port) < 1
            if waiting:
                if port not in self.queue:
                    self.queue.insert(0, port)
            elif port in self.queue:
                self.queue.remove(port)

    def get_action(self, env) -> np.ndarray:                                         # :54-95
        want = float(env.scenario.setpoint[env.current_step]) * 1000 / self.average_power
        self._refresh(env)
        k = min(int(np.ceil(want)), len(self.queue))
        served, self.queue = self.queue[:k], self.queue[k:]
        self.queue.extend(served)
        act = np.zeros(env.topo.P)
        for i, port in enumerate(served):
            act[port] = 1 / self.ports_per_cs
            if i == len(served) - 1 and want < len(served):
                act[port] = want - i
        return act


class OracleChargeAsLateAsPossible:
    """heuristics.py:98-150: full power from the last step that still fills the battery by departure."""

    def get_action(self, env) -> np.ndarray:
        tp, sc = env.topo, env.scenario
        act = np.zeros(tp.P)
        for port, (c, _) in enumerate(_ports(env)):
            s = int(env.arr["port_session"][port])
            if s < 0:
                continue
            cs_kw = float(tp.cs_imax[c]) * float(tp.cs_voltage[c]) * math.sqrt(int(tp.cs_phases[c])) / 1000   # :120-121
            kw = min(cs_kw, float(sc.sessions["pmax_ac"][s]))                                              # :123-124
            soc = _soc(env, port)
            steps = math.ceil((1 - soc) / (kw * tp.timescale / 60 / float(sc.sessions["B"][s])))            # :127-128
            if soc < 1 and int(sc.sessions["t_dep"][s]) - steps <= env.current_step:                        # :130-132
                act[port] = 1
        return act
