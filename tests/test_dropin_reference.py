"""Proof of the drop-in claim (north_star: "the baselines.heuristics agents drop in unchanged"), on the CPU.

The UNMODIFIED reference classes -- `ev2gym.baselines.heuristics.RoundRobin`, `ChargeAsFastAsPossibleToDesiredCapacity`,
`ChargeAsLateAsPossibleToDesiredCapacity` -- and the reference's own reward / state function OBJECTS
(`ev2gym.rl_agent.reward`, `.state`) drive `EV2GymB200(config_file=..., scenario_source="reference")`, the documented
default construction, and the episode is compared step by step with `ev2gym.models.ev2gym_env.EV2Gym` itself built
from the same config and seed: the action vector each agent computes from the env's attributes, observation, reward,
done, info["action_mask"], the histories (`cs_power[C,T]`, `tr_overload[Tr,T]`, `current_power_usage[T]`,
`charge_power_potential[T]`), `sim_date`, and the end-of-episode statistics dict.

Two plugin modes per case:
  fused   the stock callables are recognised by `__name__` and run inside the CUDA kernel
  python  the same callables wrapped so that they are NOT recognised: the facade then calls them like the reference does
          (ev2gym_env.py:563-565, 579-586) on the attribute-compatible views of compat.py

The container has no GPU, so the engine under the facade is the SIMT-emulated build of the same CUDA sources
(tests/simt_emu; `ev2gym_b200.env._ENGINE_CLS` is pointed at its adapter).  The reference is imported read-only from
/root/reference through the stub packages of oracle/refshim; where it is absent (the GPU box) the module is skipped.
tests/test_gpu_facade.py runs the facade on the real library.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ev2gym")), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def ref():
    """The reference modules, imported from /root/reference (cwd there: shipped configs use relative data paths)."""
    import types
    import warnings
    warnings.filterwarnings("ignore")
    old_cwd, old_path = os.getcwd(), list(sys.path)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    sys.dont_write_bytecode = True
    os.chdir(REF)
    from ev2gym.baselines import heuristics
    from ev2gym.models.ev2gym_env import EV2Gym
    from ev2gym.rl_agent import reward, state
    import emu_engine
    import ev2gym_b200.env as b200env
    emu_engine.build()
    saved = b200env._ENGINE_CLS
    b200env._ENGINE_CLS = emu_engine.EmuTorchEngine
    yield types.SimpleNamespace(EV2Gym=EV2Gym, heuristics=heuristics, reward=reward, state=state, env=b200env)
    b200env._ENGINE_CLS = saved
    os.chdir(old_cwd)
    sys.path[:] = old_path


def _config(base, overrides):
    import yaml
    cfg = yaml.safe_load(open(f"{REF}/ev2gym/example_config_files/{base}.yaml"))
    cfg.update(overrides)
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    return f.name


def _agent(ref, name, env):
    cls = getattr(ref.heuristics, name)
    return cls(env) if name == "RoundRobin" else cls()


def _unrecognised(fn):
    """The same callable under a name the facade does not know: forces the Python plugin path."""
    def plugin(*a, **k):
        return fn(*a, **k)
    return plugin


CASES = [
    # base config, overrides, seed, agent class, state fn, reward fn
    ("PublicPST", {"number_of_charging_stations": 12}, 3, "RoundRobin", "PublicPST", "SquaredTrackingErrorReward"),
    ("V2GProfitPlusLoads", {"number_of_charging_stations": 10}, 5, "ChargeAsFastAsPossibleToDesiredCapacity",
     "V2G_profit_max_loads", "ProfitMax_TrPenalty_UserIncentives"),
    ("V2GProfitMax", {"number_of_charging_stations": 8}, 9, "ChargeAsLateAsPossibleToDesiredCapacity",
     "V2G_profit_max", "profit_maximization"),
    ("V2GProfitPlusLoads", {"number_of_charging_stations": 6, "number_of_ports_per_cs": 2, "number_of_transformers": 2,
      "power_setpoint_enabled": True}, 11,
     "RoundRobin", "V2G_profit_max_loads", "ProfitMax_TrPenalty_UserIncentives"),
]


@pytest.mark.parametrize("mode", ["fused", "python"])
@pytest.mark.parametrize("base,overrides,seed,agent,st,rw", CASES)
def test_reference_agents_and_plugins_drop_in(ref, base, overrides, seed, agent, st, rw, mode):
    state_fn, reward_fn = getattr(ref.state, st), getattr(ref.reward, rw)
    path = _config(base, overrides)
    try:
        gold = ref.EV2Gym(config_file=path, seed=seed, state_function=state_fn, reward_function=reward_fn)
        mine = ref.env.EV2GymB200(config_file=path, seed=seed,
                                  state_function=state_fn if mode == "fused" else _unrecognised(state_fn),
                                  reward_function=reward_fn if mode == "fused" else _unrecognised(reward_fn))
    finally:
        os.unlink(path)
    assert mine._fused_state == (mode == "fused") and mine._fused_reward == (mode == "fused")
    obs_g, _ = gold.reset(seed=seed)
    obs_m, _ = mine.reset(seed=seed)
    assert mine.sim_date == gold.sim_date
    assert mine.number_of_ports == gold.number_of_ports and mine.simulation_length == gold.simulation_length
    tol = dict(rtol=1e-5, atol=1e-5) if mode == "fused" else dict(rtol=1e-9, atol=1e-9)   # fused observations are float32
    assert np.allclose(obs_m, obs_g, **tol)
    ag, am = _agent(ref, agent, gold), _agent(ref, agent, mine)
    T = gold.simulation_length
    nonzero_actions = 0
    for t in range(T):
        a_g = np.asarray(ag.get_action(gold), dtype=np.float64)
        a_m = np.asarray(am.get_action(mine), dtype=np.float64)
        assert np.array_equal(a_m, a_g), (t, "the agent computes a different action from the facade's attributes")
        nonzero_actions += int(np.count_nonzero(a_g))
        og, rg, dg, _, ig = gold.step(a_g)
        om, rm, dm, _, im = mine.step(a_m)
        assert np.array_equal(a_m, a_g), (t, "in-place zeroing of empty-port actions")        # ev_charger.py:137-140
        assert np.allclose(om, og, **tol), (t, "observation")
        assert rm == pytest.approx(rg, rel=1e-9, abs=1e-9), (t, "reward")
        assert dm == dg
        assert np.array_equal(np.asarray(im["action_mask"]), np.asarray(ig["action_mask"])), (t, "action mask")
        assert mine.current_step == gold.current_step and mine.sim_date == gold.sim_date
        for p, (cg, cm) in enumerate(zip(gold.charging_stations, mine.charging_stations)):
            for j in range(cg.n_ports):
                eg, em = cg.evs_connected[j], cm.evs_connected[j]
                assert (eg is None) == (em is None), (t, p, j)
                if eg is not None:
                    assert em.current_capacity == eg.current_capacity, (t, p, j, "battery level")   # bit exact
                    assert em.time_of_departure == eg.time_of_departure and em.desired_capacity == eg.desired_capacity
    assert nonzero_actions > 0, "the case never asked any EV to charge: it proves nothing"
    assert dg and dm
    assert np.allclose(mine.cs_power, gold.cs_power, rtol=1e-5, atol=1e-6)                 # [C, T] histories
    assert np.allclose(mine.cs_current, gold.cs_current, rtol=1e-5, atol=1e-6)
    assert np.allclose(mine.tr_overload, gold.tr_overload, rtol=1e-9, atol=1e-9)            # [Tr, T]
    assert np.allclose(mine.current_power_usage, gold.current_power_usage, rtol=1e-9, atol=1e-9)
    assert np.allclose(mine.charge_power_potential, gold.charge_power_potential, rtol=1e-9, atol=1e-9)
    assert mine.total_reward == pytest.approx(gold.total_reward, rel=1e-9, abs=1e-9)
    for k, v in ig.items():                                                                  # get_statistics(env)
        if k == "action_mask" or not np.isscalar(v):
            continue
        assert k in im, k
        if isinstance(v, float) and np.isnan(v):
            assert np.isnan(im[k]), k
        else:
            assert im[k] == pytest.approx(v, rel=1e-9, abs=1e-9), k
    mine.close()
