"""GPU: the drop-in facade (EV2GymB200) replays reference episodes: gym 5-tuple, spaces, stats, the
in-place zeroing of empty-port actions, custom Python plugin functions on the compat views."""
import math

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = ["c1_afap_s42", "pst25_uniform_s3", "loads_c20n2tr3_mixed_s11", "profitmax_c25_mixed_s9",
         "mincur_c6n2_mixed_s8", "grid_c40_uniform_s3"]


def _env(name, **kw):
    from ev2gym_b200.env import EV2GymB200
    from ev2gym_b200.scenario import ScenarioPack
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    kw.setdefault("state_function", str(tr["state_fn"]))
    kw.setdefault("reward_function", str(tr["reward_fn"]))
    return EV2GymB200(scenario_source=pack, **kw), tr


@pytest.mark.parametrize("name", CASES)
def test_facade_replays_reference_episode(name):
    env, tr = _env(name)
    obs0, info = env.reset()
    assert info == {} and np.allclose(obs0, tr["obs0"], rtol=1e-5, atol=1e-5)
    assert env.action_space.shape == (env.number_of_ports,) and env.observation_space.shape == obs0.shape
    T = tr["reward"].shape[0]
    for t in range(T):
        a = tr["actions"][t].copy()
        obs, r, done, trunc, info = env.step(a)
        assert np.array_equal(a, tr["actions_eff"][t]), "empty-port actions must be zeroed in the caller's array"
        assert np.allclose(obs, tr["obs"][t], rtol=1e-5, atol=1e-5), t
        assert r == pytest.approx(tr["reward"][t], rel=1e-9, abs=1e-9)
        assert done == bool(tr["done"][t]) and trunc is False
        assert np.array_equal(info["action_mask"], tr["action_mask"][t])
    assert info["total_ev_served"] == int(tr["stat_total_ev_served"])
    for k in ("total_profits", "total_energy_charged", "total_energy_discharged", "total_transformer_overload",
              "tracking_error", "energy_tracking_error", "power_tracker_violation", "energy_user_satisfaction",
              "std_energy_user_satisfaction", "min_energy_user_satisfaction", "battery_degradation",
              "battery_degradation_calendar", "battery_degradation_cycling", "total_reward",
              "total_steps_min_emergency_battery_capacity_violation"):
        assert info[k] == pytest.approx(float(tr["stat_" + k]), rel=1e-9, abs=1e-9), k
    if not math.isnan(float(tr["stat_average_user_satisfaction"])):
        assert info["average_user_satisfaction"] == pytest.approx(float(tr["stat_average_user_satisfaction"]), rel=1e-9)
    assert np.allclose(env.current_power_usage, tr["usage"], rtol=1e-12, atol=1e-12)
    assert np.allclose(env.charge_power_potential, tr["potential"], rtol=1e-12, atol=1e-12)
    with pytest.raises(AssertionError):
        env.step(tr["actions"][0].copy())
    env.close()


def my_reward(env, total_costs, user_satisfaction_list, *args):
    """Same arithmetic as the stock ProfitMax_TrPenalty_UserIncentives, under a name the engine does not know,
    so it runs through the attribute-compatible views exactly like a user-defined plugin."""
    reward = total_costs
    for tr in env.transformers:
        reward -= 100 * tr.get_how_overloaded()
    for score in user_satisfaction_list:
        reward -= 100 * math.exp(-10 * score)
    return reward


def my_state(env, *args):
    state = [env.current_step, env.current_power_usage[env.current_step - 1]]
    prices = abs(env.charge_prices[0, env.current_step:env.current_step + 20])
    if len(prices) < 20:
        prices = np.append(prices, np.zeros(20 - len(prices)))
    state.append(prices)
    for tr in env.transformers:
        loads, pv = tr.get_load_pv_forecast(step=env.current_step, horizon=20)
        state.append(loads - pv)
        state.append(tr.get_power_limits(step=env.current_step, horizon=20))
        for cs in env.charging_stations:
            if cs.connected_transformer == tr.id:
                for ev in cs.evs_connected:
                    state.append([ev.get_soc(), ev.time_of_departure - env.current_step] if ev is not None
                                 else np.zeros(2))
    return np.array(np.hstack(state))


def test_custom_python_plugins_on_compat_views():
    env, tr = _env("loads_c20n2tr3_mixed_s11", state_function=my_state, reward_function=my_reward)
    obs0, _ = env.reset()
    assert np.allclose(obs0, tr["obs0"], rtol=1e-12, atol=1e-12)
    for t in range(tr["reward"].shape[0]):
        obs, r, done, _, info = env.step(list(tr["actions"][t]))        # a python list, like the tutorials pass
        assert np.allclose(obs, tr["obs"][t], rtol=1e-9, atol=1e-9), t  # float64 end to end on this path
        assert r == pytest.approx(tr["reward"][t], rel=1e-9, abs=1e-9)
        caps = [ev.current_capacity for cs in env.charging_stations for ev in cs.evs_connected if ev is not None]
        assert caps == [c for c in tr["cap"][t] if not np.isnan(c)]      # bit exact battery levels
        assert len(env.departing_evs) == tr["n_departed"][t]
    assert env.total_reward == pytest.approx(float(tr["total_reward"]), rel=1e-9)
    env.close()


def test_vec_env_auto_reset():
    import torch
    from ev2gym_b200.env import EV2GymB200Vec
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    topo = Topology.uniform(C=10, n_ports=1, Tr=1, T=24)
    vec = EV2GymB200Vec(topo, sample_bank(topo, 4, seed=2, min_stay=4), num_envs=32)
    obs = vec.reset()
    assert obs.shape == (32, vec.obs_dim)
    n_done = 0
    for t in range(50):
        obs, r, done, info = vec.step(torch.rand(32, topo.P, device="cuda") * 2 - 1)
        n_done += int(done.sum())
        if done.any():
            assert float(obs[done][:, 0].max()) == 0.0          # already the first observation of the next episode
    assert n_done == 64 and int(vec.state_tensors()["env_step"][0]) == 2


def test_vec_env_histories_replay_reference_episode():
    """EV2GymB200Vec(histories=True) on a recorded reference episode: after the episode the device-side histories equal
    the reference's env.cs_power[C,T] / cs_current / tr_overload[Tr,T] / current_power_usage[T] (fixtures recorded by
    tools/make_golden.py), the mask of a finished env is cleared by the auto reset, sim_minutes follows _step_date."""
    import torch
    from ev2gym_b200.env import EV2GymB200Vec
    from ev2gym_b200.scenario import ScenarioPack
    name = "loads_c20n2tr3_mixed_s11"
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    E = 3
    vec = EV2GymB200Vec(pack.topo, pack.scenarios, E, state_function=str(tr["state_fn"]), reward_function=str(tr["reward_fn"]),
                        histories=True)
    vec.reset()
    T = tr["reward"].shape[0]
    for t in range(T):
        a = torch.tensor(np.tile(tr["actions"][t], (E, 1)), device="cuda")
        assert int(vec.sim_minutes()[0]) == t * pack.topo.timescale
        obs, r, done, info = vec.step(a)
    assert bool(done.all()) and int(info["action_mask"].sum()) == 0
    h = {k: v.cpu().numpy() for k, v in vec.histories().items()}
    for e in range(E):
        assert np.allclose(h["cs_power"][e], tr["cs_power"].T, rtol=1e-5, atol=1e-6)
        assert np.allclose(h["cs_current"][e], tr["cs_current"].T, rtol=1e-5, atol=1e-6)
        assert np.allclose(h["tr_overload"][e], tr["tr_overload"].T, rtol=1e-9, atol=1e-9)
        assert np.allclose(h["current_power_usage"][e], tr["usage"][:T], rtol=1e-9, atol=1e-9)


def test_sb3_style_vec_env_replays_reference_episode():
    """EV2GymB200SB3Vec (numpy VecEnv convention: step_async / step_wait, auto reset, terminal_observation) on a
    recorded ChargeAsFastAsPossible episode: per-step obs / reward / done, the terminal observation and the episode
    return equal the reference's; after the last step every env already shows the first observation of its next episode."""
    from ev2gym_b200.env import EV2GymB200SB3Vec
    from ev2gym_b200.scenario import ScenarioPack
    name = "c1_afap_s42"
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    E = 5
    venv = EV2GymB200SB3Vec(pack.topo, pack.scenarios, E, state_function=str(tr["state_fn"]),
                            reward_function=str(tr["reward_fn"]))
    assert venv.num_envs == E and venv.action_space.shape == (pack.topo.P,)
    obs = venv.reset()
    assert obs.shape == (E,) + venv.observation_space.shape and obs.dtype == np.float32
    assert np.allclose(obs[0], tr["obs0"], rtol=1e-5, atol=1e-5)
    T = tr["reward"].shape[0]
    for t in range(T):
        venv.step_async(np.ones((E, pack.topo.P), dtype=np.float32))
        obs, rew, done, infos = venv.step_wait()
        assert rew.dtype == np.float32 and done.dtype == bool and len(infos) == E
        assert np.allclose(rew, tr["reward"][t], rtol=1e-5, atol=1e-5), t
        assert bool(done.all()) == bool(tr["done"][t])
        if t < T - 1:
            assert np.allclose(obs[E - 1], tr["obs"][t], rtol=1e-5, atol=1e-5), t
            assert infos[0] == {}
    for e in (0, E - 1):
        assert np.allclose(infos[e]["terminal_observation"], tr["obs"][T - 1], rtol=1e-5, atol=1e-5)
        assert infos[e]["episode"]["l"] == T
        assert infos[e]["episode"]["r"] == pytest.approx(float(tr["total_reward"]), rel=1e-9)
        assert np.allclose(obs[e], tr["obs0"], rtol=1e-5, atol=1e-5)          # bank of one scenario: the same episode restarts
    obs2, rew2, done2, _ = venv.step(np.ones((E, pack.topo.P), dtype=np.float32))   # and it runs on
    assert np.allclose(rew2, tr["reward"][0], rtol=1e-5, atol=1e-5) and not done2.any()
    assert venv.get_attr("simulation_length") == [T] * E
    venv.close()
