"""Drop-in facades over the CUDA step engine.

`EV2GymB200` mirrors `ev2gym.models.ev2gym_env.EV2Gym` (constructor keywords, `reset`, `step`,
`set_reward_function`, `set_cost_function`, spaces, and the attributes plugin functions read;
reference: ev2gym/models/ev2gym_env.py:38-56, 243-331, 333-447, 567-577) for ONE env whose
`step()` is the fused kernel.  Stock state / reward functions (identified by `__name__`) run fused
on the device; any other callable is invoked exactly like the reference does, on attribute-compatible
views (compat.py) rebuilt from the device state after each step.

`EV2GymB200Vec` is the batched surface for RL: E replicas, torch tensors in and out, device-side
auto-reset; nothing crosses PCIe unless the caller asks.

Scenario generation (the reference's `reset()` loaders) is out of scope for the GPU path
(SURVEY.md section 2 rows 5, 7): it stays on the host.  `scenario_source` selects it:
  "reference"  the reference package itself (must be importable) builds each episode; it is only
               ever reset(), never stepped
  ScenarioPack / list[Scenario]   episodes exported earlier (tools/make_golden.py --packs)
  "synthetic"  the numpy sampler of ev2gym_b200.synthetic (no reference data needed)
"""
from __future__ import annotations

import datetime
import math
from typing import Callable, List, Optional, Sequence, Union

import numpy as np

from .compat import ChargerView, EVView, TransformerView
from .engine import REWARD_KINDS, STATE_KINDS, BatchedEngine, EngineError, _fn_name
from .scenario import Scenario, ScenarioPack, Topology, assign_ports

# The engine class the facades instantiate.  The product has exactly one: BatchedEngine (CUDA, raises without a GPU).
# tests/test_dropin_reference.py points this at an adapter over the SIMT emulator (tests/simt_emu) so that the facade's
# HOST logic -- views, bookkeeping, plugin calls -- can be checked against the reference in a container without a GPU.
_ENGINE_CLS = BatchedEngine

_FACADE_OUTPUTS_GRID = ("node_voltage",)
_FACADE_OUTPUTS = ("reward", "status", "obs", "cs_power", "cs_current", "tr_power", "tr_overload", "total_costs",
                   "action_mask", "dep_sat", "dep_cap", "port_energy")


class Box:
    """Duck-typed gymnasium.spaces.Box (gymnasium is not a dependency of this package)."""

    def __init__(self, low, high, dtype=np.float64):
        self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        self.shape, self.dtype = self.low.shape, np.dtype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _import_reference():
    try:
        from ev2gym.models.ev2gym_env import EV2Gym  # noqa: WPS433 (optional dependency)
        return EV2Gym
    except Exception as exc:  # pragma: no cover - depends on the user's environment
        raise EngineError("scenario_source='reference' needs the `ev2gym` package importable "
                          f"({type(exc).__name__}: {exc}); pass a ScenarioPack or 'synthetic' instead") from exc


class EV2GymB200:
    metadata = {}

    def __init__(self, config_file=None, load_from_replay_path=None, replay_save_path="./replay/",
                 generate_rnd_game=True, seed=None, save_replay=False, save_plots=False,
                 state_function="PublicPST", reward_function="SquaredTrackingErrorReward", cost_function=None,
                 eval_mode="Normal", lightweight_plots=False, empty_ports_at_end_of_simulation=True,
                 extra_sim_name=None, verbose=False, render_mode=None,
                 # --- additions (keyword only in spirit; never required) ---
                 scenario_source: Union[str, ScenarioPack, Sequence[Scenario]] = "reference",
                 topology: Optional[Topology] = None, device: int = 0):
        if save_replay or save_plots or render_mode:
            raise NotImplementedError("replay / plots / rendering are out of scope of the GPU step path")
        self.state_function, self.reward_function, self.cost_function = state_function, reward_function, cost_function
        self.verbose = verbose
        self._ref = None
        self._scn_iter = 0
        self.seed = seed
        if isinstance(scenario_source, str) and scenario_source == "reference":
            from .reference_export import scenario_from_env, topology_from_env
            ref_state = state_function if callable(state_function) else None
            ref_reward = reward_function if callable(reward_function) else None
            kw = dict(config_file=config_file, load_from_replay_path=load_from_replay_path, seed=seed,
                      generate_rnd_game=generate_rnd_game, empty_ports_at_end_of_simulation=empty_ports_at_end_of_simulation,
                      extra_sim_name=extra_sim_name, eval_mode=eval_mode, lightweight_plots=True)
            if ref_state is not None:
                kw["state_function"] = ref_state
            if ref_reward is not None:
                kw["reward_function"] = ref_reward
            self._ref = _import_reference()(**kw)
            self._export = scenario_from_env
            self.topo = topology_from_env(self._ref)
            self.config = self._ref.config
            self._scenarios = None
        else:
            if isinstance(scenario_source, str) and scenario_source == "synthetic":
                from .synthetic import sample_bank
                if topology is None:
                    raise EngineError("scenario_source='synthetic' needs topology=Topology(...)")
                self.topo = topology
                self._scenarios = sample_bank(topology, 16, seed=0 if seed is None else seed)
            elif isinstance(scenario_source, ScenarioPack):
                self.topo, self._scenarios = scenario_source.topo, list(scenario_source.scenarios)
            else:
                if topology is None:
                    raise EngineError("a list of scenarios needs topology=Topology(...)")
                self.topo, self._scenarios = topology, list(scenario_source)
            self.config = {"v2g_enabled": self.topo.v2g_enabled, "timescale": self.topo.timescale,
                           "simulation_length": self.topo.T}
        topo = self.topo
        self.simulation_length, self.timescale = topo.T, topo.timescale
        self.cs, self.number_of_ports = topo.C, topo.P
        self.number_of_ports_per_cs = int(topo.cs_n_ports[0])
        self.number_of_transformers = topo.Tr
        self.cs_transformers = [int(x) for x in topo.cs_tr]
        self.simulate_grid = False
        self._fused_state = _fn_name(state_function) in STATE_KINDS and _fn_name(state_function) is not None
        self._fused_reward = _fn_name(reward_function) in REWARD_KINDS and _fn_name(reward_function) is not None
        self._engine = _ENGINE_CLS(topo, 1, reward=reward_function if self._fused_reward else None,
                                     state=state_function if self._fused_state else None, device=device,
                                     outputs=_FACADE_OUTPUTS + (_FACADE_OUTPUTS_GRID if topo.n_bus else ()), stats=True)
        self._port_off = topo.cs_port_off
        self.done = False
        high = np.ones(self.number_of_ports)
        lows = -1 * np.ones(self.number_of_ports) if topo.v2g_enabled else np.zeros(self.number_of_ports)
        self.action_space = Box(lows, high, np.float64)                      # ev2gym_env.py:225-231
        obs = self.reset(seed=seed)[0]
        inf = np.inf * np.ones(len(obs))
        self.observation_space = Box(-inf, inf, np.float64)                  # ev2gym_env.py:233-238
        self.observation_mask = np.zeros(self.number_of_ports)

    # ------------------------------------------------------------------------------------------
    def _next_scenario(self, seed) -> Scenario:
        if self._ref is not None:
            self._ref.reset(seed=seed)
            self.sim_date = self._ref.sim_date                       # ev2gym_env.py:262-287 (stepped below, :558)
            self.sim_starting_date = getattr(self._ref, "sim_starting_date", self.sim_date)
            return self._export(self._ref)
        sc = self._scenarios[self._scn_iter % len(self._scenarios)]
        self._scn_iter += 1
        self.sim_date = self.sim_starting_date = getattr(sc, "sim_date", None)
        return sc

    def reset(self, seed=None, options=None, **kwargs):
        """Samples a new episode on the host and resets the device state (ev2gym_env.py:243-331)."""
        topo = self.topo
        sc = self._next_scenario(seed)
        self._sc = sc
        self._engine.load_scenarios([sc])
        obs = self._engine.reset()
        s = sc.sessions
        port = assign_ports(topo, s["t_arr"], s["t_dep"], s["loc"])
        self._port_sessions: List[List[int]] = [[] for _ in range(topo.P)]
        for i, p in enumerate(port):
            self._port_sessions[p].append(i)
        T = topo.T
        self.current_step = 0
        self.done = False
        self.stats = None
        self.total_reward = 0.0
        self.total_evs_spawned = 0
        self.current_evs_parked = 0
        self.current_ev_departed = self.current_ev_arrived = 0
        self.power_setpoints = sc.setpoint
        self.charge_prices = np.tile(sc.charge_price, (topo.C, 1))
        self.discharge_prices = np.tile(sc.discharge_price, (topo.C, 1))
        self.current_power_usage = np.zeros(T)
        self.charge_power_potential = np.zeros(T)
        self.cs_power, self.cs_current = np.zeros((topo.C, T)), np.zeros((topo.C, T))
        self.tr_overload = np.zeros((topo.Tr, T))
        self.tr_inflexible_loads, self.tr_solar_power = sc.tr_infl.copy(), sc.tr_solar.copy()
        nb = topo.n_bus + 1 if topo.n_bus else 34
        self.simulate_grid = topo.n_bus > 0
        self.node_active_power = np.zeros((nb, T))                      # ev2gym_env.py:321-327, 395-396
        self.node_reactive_power = np.zeros((nb, T))
        self.node_voltage = np.zeros((nb, T))
        if self.simulate_grid:
            self.node_active_power[1:, 0], self.node_reactive_power[1:, 0] = sc.grid_active[0], sc.grid_reactive[0]
        self.departing_evs: List[EVView] = []
        self.EVs: List[EVView] = []
        self.EVs_profiles = [EVView(sc, i, 0, s["cap0"][i], 0.0, 0.0, 0.0, topo.timescale) for i in range(sc.n_sessions)]
        self.charging_stations = [ChargerView(topo, c) for c in range(topo.C)]
        self.transformers = [TransformerView(topo, sc, k) for k in range(topo.Tr)]
        self._mask = np.zeros(topo.P)
        if self._fused_state:
            state = obs[0].cpu().numpy().astype(np.float64)
        else:
            state = np.asarray(self.state_function(self))
        return state, {}

    # ------------------------------------------------------------------------------------------
    def _refresh_views(self, out, cap, exch, hot, t_done):
        topo, sc = self.topo, self._sc
        dec = BatchedEngine.decode_hot(hot)
        self.departing_evs = []
        sat_list = []
        for c, cs in enumerate(self.charging_stations):
            lo = int(self._port_off[c])
            cs.current_power_output = float(out["cs_power"][c])
            cs.current_total_amps = float(out["cs_current"][c])
            cs.current_step = t_done + 1
            cs.current_charge_price = float(sc.charge_price[t_done])
            cs.current_discharge_price = float(sc.discharge_price[t_done])
            for j in range(cs.n_ports):
                p = lo + j
                prev = cs.evs_connected[j]
                e = float(out["port_energy"][p])
                if prev is not None and e != 0.0:
                    if e > 0:
                        cs.total_energy_charged += abs(e)
                    else:
                        cs.total_energy_discharged += abs(e)
                if not math.isnan(out["dep_sat"][p]):                 # ev_charger.py:209-224
                    prev.current_capacity = float(out["dep_cap"][p])
                    prev.current_energy = e
                    sat = float(out["dep_sat"][p])
                    cs.total_evs_served += 1
                    cs.total_user_satisfaction += sat
                    cs.all_user_satisfaction.append(sat)
                    sat_list.append(sat)
                    self.departing_evs.append(prev)
                    prev = None
                if out["action_mask"][p]:
                    k = int(dec["cursor"][p]) - 1
                    i = self._port_sessions[p][k]
                    if prev is None or prev._i != i:                    # spawned at the end of this step
                        prev = EVView(sc, i, j, cap[p], 0.0, 0.0, 0.0, topo.timescale)
                        self.EVs.append(prev)
                        self.current_ev_arrived += 1
                    else:
                        ph = min(cs.phases, prev.ev_phases)
                        prev.current_capacity = float(cap[p])
                        prev.total_energy_exchanged = float(exch[p])
                        prev.current_energy = e
                        prev.actual_current = e * 60 / topo.timescale * 1000 / (cs.voltage * math.sqrt(ph))
                        prev.required_energy = prev.battery_capacity - prev.current_capacity
                    cs.evs_connected[j] = prev
                else:
                    cs.evs_connected[j] = None
            cs.n_evs_connected = sum(e is not None for e in cs.evs_connected)
        for k, tr in enumerate(self.transformers):
            tr.current_step = t_done
            tr.current_power = float(out["tr_power"][k])
        return sat_list

    def step(self, actions, visualize=False):
        """One timestep on the GPU; same contract as EV2Gym.step (ev2gym_env.py:333-447)."""
        assert not self.done, "Episode is done, please reset the environment"
        import torch
        topo, eng = self.topo, self._engine
        assert len(actions) == self.number_of_ports
        a = np.asarray(actions, dtype=np.float64)
        try:                                     # the reference zeroes empty-port entries in the caller's array
            idx = np.nonzero(self._mask == 0)[0]
            for i in idx:
                actions[i] = 0
        except (TypeError, ValueError):
            pass
        t = self.current_step
        self.current_ev_departed = self.current_ev_arrived = 0
        dev_a = torch.from_numpy(np.ascontiguousarray(a)).reshape(1, -1).to(eng.dev)
        outs = eng.step(dev_a)
        st = eng.state_tensors()
        out = {k: v[0].cpu().numpy() for k, v in outs.items()}
        cap, exch = st["port_cap"][0].cpu().numpy(), st["port_exch"][0].cpu().numpy()
        hot = st["port_hot"][0].cpu().numpy()
        status = int(out["status"])
        if status & 2:                           # ev_charger.py:203-205
            raise Exception("sum of amps is higher than max charge current")
        self.current_power_usage[t] = float(st["env_usage"][0].item())
        self.cs_power[:, t], self.cs_current[:, t] = out["cs_power"], out["cs_current"]
        self.tr_overload[:, t] = out["tr_overload"]
        if self.simulate_grid:                                           # ev2gym_env.py:392-397
            self.node_voltage[:, t] = out["node_voltage"]
            self.node_active_power[1:, t] = self._sc.grid_active[t + 1]
            self.node_reactive_power[1:, t] = self._sc.grid_reactive[t + 1]
        sat_list = self._refresh_views(out, cap, exch, hot, t)
        self.current_ev_departed = len(sat_list)
        self.total_evs_spawned += self.current_ev_arrived
        self.current_step = t + 1
        if self.sim_date is not None:                                    # _step_date  ev2gym_env.py:422, 558-561
            self.sim_date = self.sim_date + datetime.timedelta(minutes=self.timescale)
        if self.current_step < self.simulation_length:
            self.charge_power_potential[self.current_step] = float(st["env_potential"][0].item())
        self.current_evs_parked += self.current_ev_arrived - self.current_ev_departed
        total_costs = float(out["total_costs"])
        invalid = int(np.count_nonzero(self._mask == 0))
        self._mask = out["action_mask"].astype(np.float64)
        if self._fused_reward:
            reward = float(out["reward"])
        else:
            reward = self.reward_function(self, total_costs, sat_list, invalid)      # ev2gym_env.py:579-586
        self.total_reward += reward
        cost = self.cost_function(self, total_costs, sat_list, invalid) if self.cost_function is not None else None
        if self._fused_state:
            obs = out["obs"].astype(np.float64)
        else:
            obs = np.asarray(self.state_function(self))
        if self.current_step >= self.simulation_length:                               # ev2gym_env.py:460-486
            self.done = True
            self.stats = self._statistics()
            self.stats["action_mask"] = self._mask.copy()
            self.cost = cost
            return obs, reward, True, False, self.stats
        return obs, reward, False, False, {"cost": cost, "action_mask": self._mask.copy()}

    def _statistics(self) -> dict:
        """get_statistics (ev2gym/utilities/utils.py:12-123), computed on the device (ev2b_episode_stats)."""
        st = {k: float(v[0]) for k, v in self._engine.episode_stats().items()}
        st["total_ev_served"] = int(st["total_ev_served"])
        st["total_steps_min_emergency_battery_capacity_violation"] = int(
            st["total_steps_min_emergency_battery_capacity_violation"])
        st["total_reward"] = self.total_reward
        st.update(saved_grid_energy=0, voltage_violation=0, voltage_violation_counter=0,
                  voltage_violation_counter_per_step=0)                              # utils.py:108-112
        return st

    def set_cost_function(self, cost_function):
        self.cost_function = cost_function

    def set_reward_function(self, reward_function):
        if _fn_name(reward_function) in REWARD_KINDS and _fn_name(reward_function) != self._engine.reward_name:
            raise NotImplementedError("switching between fused rewards needs a new env; pass a plain callable instead")
        self.reward_function = reward_function
        self._fused_reward = _fn_name(reward_function) == self._engine.reward_name and self._engine.reward_name is not None

    def close(self):
        self._engine.close()


class EV2GymB200Vec:
    """E env replicas stepped by one kernel launch; torch in / torch out; device-side auto reset."""

    def __init__(self, topo: Topology, scenarios: Sequence[Scenario], num_envs: int, state_function="V2G_profit_max",
                 reward_function="profit_maximization", device: int = 0, auto_reset: bool = True, rank: int = 0,
                 histories: bool = False):
        """histories=True keeps, per env, what the reference keeps for plots / statistics / custom plugins
        (init_statistic_variables, utils.py:794-861): `self.histories()` -> cs_power / cs_current [E,C,T],
        tr_overload [E,Tr,T], current_power_usage [E,T], filled row by row as the episode runs; `self.sim_step` is the
        per-env step counter, `self.sim_minutes()` the minutes since each env's sim_date (`_step_date`, ev2gym_env.py:558)."""
        self.topo, self.num_envs, self.auto_reset = topo, num_envs, auto_reset
        hist = ("hist_cs_power", "hist_cs_current", "hist_tr_overload", "hist_usage") if histories else ()
        self.engine = _ENGINE_CLS(topo, num_envs, reward=reward_function, state=state_function, device=device,
                                  outputs=("reward", "status", "obs", "action_mask") + hist)
        self.engine.load_scenarios(scenarios)
        self._first = [(rank * num_envs + e) % len(scenarios) for e in range(num_envs)]
        self.obs_dim, self.n_actions = self.engine.D, topo.P
        self.action_low = -1.0 if topo.v2g_enabled else 0.0

    def reset(self):
        return self.engine.reset(scn_ids=self._first)

    def step(self, actions):
        """actions: cuda tensor [E,P] (fp32/fp64).  Returns (obs, reward, done, info) tensors; with auto_reset
        the returned obs rows of finished envs are already the first observation of their next episode (and their
        action-mask rows are all zero: no EV is connected at t = 0).

        `obs` and `info["action_mask"]` are the engine's LIVE buffers: the next step() updates them in place (only the
        entries that changed are rewritten, include/ev2b.h), so read them -- or .clone() them -- before stepping again,
        and do not write into them (e.g. in-place observation normalisation): normalise a copy.  `reward` and `done`
        are fresh tensors."""
        out = self.engine.step(actions)
        done = (out["status"] & 1).bool()
        reward = out["reward"].clone()
        info = {"action_mask": out["action_mask"]}
        if self.auto_reset:
            info["terminal_obs_overwritten"] = True
            self.engine.reset_done()
            out["action_mask"].masked_fill_(done.unsqueeze(1), 0)     # the terminal mask of the finished episode
        return out["obs"], reward, done, info

    def state_tensors(self):
        return self.engine.state_tensors()

    def histories(self):
        return self.engine.histories()

    @property
    def sim_step(self):
        return self.engine.state_tensors()["env_step"]

    def sim_minutes(self):
        """Minutes elapsed since each env's sim_date (the reference steps a datetime, ev2gym_env.py:558-561)."""
        return self.sim_step * self.topo.timescale


class EV2GymB200SB3Vec:
    """The same batch behind the VecEnv call convention of stable-baselines3 (duck-typed: SB3 is not imported): numpy in /
    numpy out, `step_async` + `step_wait`, auto-reset with `terminal_observation`, one info dict per env.  This is the
    surface the reference's SB3 scripts drive through `DummyVecEnv` / `SubprocVecEnv` around E separate `EV2Gym`
    processes (train_stable_baselines.py:139-175); here all E envs are one `ev2b_step_host` call with pinned buffers.

    Differences a caller can see: observations are float32 (SB3 casts to float32 anyway); `get_attr` / `env_method`
    serve only what is meaningful for a batch (`simulation_length`, `number_of_ports`, ...)."""

    def __init__(self, topo: Topology, scenarios: Sequence[Scenario], num_envs: int, state_function="V2G_profit_max",
                 reward_function="profit_maximization", device: int = 0, rank: int = 0):
        import torch
        self.topo, self.num_envs = topo, int(num_envs)
        self.engine = BatchedEngine(topo, num_envs, reward=reward_function, state=state_function, device=device,
                                    outputs=("reward", "status", "obs"))
        self.engine.load_scenarios(scenarios)
        self._first = [(rank * num_envs + e) % len(scenarios) for e in range(num_envs)]
        E, P, D = self.num_envs, topo.P, self.engine.D
        low = -1.0 if topo.v2g_enabled else 0.0
        self.action_space = Box(low * np.ones(P), np.ones(P), dtype=np.float32)                       # ev2gym_env.py:225-231
        self.observation_space = Box(-np.inf * np.ones(D), np.inf * np.ones(D), dtype=np.float32)
        self.render_mode = None
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        self._keep = [pin((E, P), torch.float32), pin((E,), torch.float64), pin((E,), torch.int32), pin((E, D), torch.float32)]
        self._act, self._rew, self._st, self._obs = (t.numpy() for t in self._keep)
        self._st = self._st.view(np.uint32)
        self._ep_ret, self._ep_len = np.zeros(E), np.zeros(E, dtype=np.int64)
        self._pending = False

    def reset(self):
        self._ep_ret[:], self._ep_len[:] = 0.0, 0
        return self.engine.reset(scn_ids=self._first).cpu().numpy().copy()

    def step_async(self, actions):
        a = np.asarray(actions)
        if a.shape != self._act.shape:
            raise ValueError(f"actions must have shape {self._act.shape}")
        self._act[...] = a                           # SB3 hands over float32 arrays; the kernel widens them to float64
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_async must be called first"
        self._pending = False
        self.engine.step_host(self._act, self._rew, self._st, self._obs)
        obs, rew = self._obs.copy(), self._rew.astype(np.float32)
        if (self._st & 2).any():
            raise Exception("sum of amps is higher than max charge current")         # ev_charger.py:203-205
        done = (self._st & 1).astype(bool)
        self._ep_ret += self._rew
        self._ep_len += 1
        infos = [{} for _ in range(self.num_envs)]
        if done.any():
            idx = np.nonzero(done)[0]
            first = self.engine.reset_done().index_select(0, self.engine.torch.as_tensor(idx, device=self.engine.dev)).cpu().numpy()
            for k, e in enumerate(idx):
                infos[e] = {"terminal_observation": obs[e].copy(), "TimeLimit.truncated": False,
                            "episode": {"r": float(self._ep_ret[e]), "l": int(self._ep_len[e])}}
                obs[e] = first[k]
            self._ep_ret[idx], self._ep_len[idx] = 0.0, 0
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.engine.close()

    def seed(self, seed=None):
        return [None] * self.num_envs                # step() consumes no randomness; scenarios are fixed by the bank

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_attr(self, name, indices=None):
        vals = {"simulation_length": self.topo.T, "number_of_ports": self.topo.P, "cs": self.topo.C,
                "number_of_transformers": self.topo.Tr, "timescale": self.topo.timescale, "render_mode": None}
        if name not in vals:
            raise AttributeError(f"{name!r} is not available on the batched env")
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [vals[name]] * n

    def set_attr(self, name, value, indices=None):
        raise AttributeError("the batched env has no per-env Python attributes to set")

    def env_method(self, method_name, *args, indices=None, **kwargs):
        raise AttributeError(f"{method_name!r}: per-env Python methods do not exist on the batched env")
