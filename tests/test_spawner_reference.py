"""Device-side EV_spawner (ev2gym_b200/csrc/ev2b_spawn.cuh, spawn_sessions_kernel) against the REFERENCE's own
`EV_spawner` / `spawn_single_EV` (ev2gym/utilities/utils.py:477-557, 177-345) on the same random numbers.

tests/test_spawn.py compares distributions (the reference consumes numpy's global stream, the device a counter-based
generator).  Here the unmodified reference functions run with numpy's `rand / normal / randint / choice` replaced by
functions that return the DEVICE generator's draws for the (scenario, spawner port, step, draw index) the reference is
at -- the keying of ev2b_spawn.cuh, restated below -- so every decision of the reference (spawn or not, required energy,
EV model, battery level at arrival, length of stay, "empty ports at the end", transition SoC, efficiencies) is taken on
the numbers the kernel saw, and the two session lists must be identical: same EVs on the same chargers at the same
steps, battery levels to 1e-12.

CPU only (the kernel runs on the SIMT emulator); skipped where /root/reference is absent.
"""
import datetime
import math
import os
import sys
import tempfile
import types

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ev2gym")), reason="needs the reference checkout")

M64 = (1 << 64) - 1


def _uniform(seed, counter):
    """spawn_uniform: splitmix64 of (seed, counter) -> [0, 1)."""
    z = (seed + (counter + 1) * 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return float(z >> 11) * (1.0 / 9007199254740992.0)


def _normal(seed, c, mean, sd):
    u1, u2 = 1.0 - _uniform(seed, c), _uniform(seed, c + 1)
    return mean + sd * math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586 * u2)


def _randint(seed, c, lo, hi):
    if hi <= lo:
        return lo
    v = lo + int(_uniform(seed, c) * float(hi - lo))
    return v if v < hi else hi - 1


@pytest.fixture(scope="module")
def ref():
    import warnings
    warnings.filterwarnings("ignore")
    old_cwd, old_path = os.getcwd(), list(sys.path)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    sys.dont_write_bytecode = True
    os.chdir(REF)
    from ev2gym.models.ev2gym_env import EV2Gym
    from ev2gym.rl_agent import reward, state
    from ev2gym.utilities import utils
    import emu_engine
    emu_engine.build()
    yield types.SimpleNamespace(EV2Gym=EV2Gym, reward=reward, state=state, utils=utils, emu=emu_engine)
    os.chdir(old_cwd)
    sys.path[:] = old_path


CASES = [
    # base config, overrides, state fn, reward fn
    ("PublicPST", {"number_of_charging_stations": 30}, "PublicPST", "SquaredTrackingErrorReward"),                   # public, homogeneous
    ("V2GProfitPlusLoads", {"number_of_charging_stations": 24, "number_of_ports_per_cs": 2, "number_of_transformers": 3},
     "V2G_profit_max_loads", "ProfitMax_TrPenalty_UserIncentives"),                                                  # heterogeneous EV models
    ("V2GProfitMax", {"number_of_charging_stations": 40}, "V2G_profit_max", "profit_maximization"),
]


@pytest.mark.parametrize("base,overrides,st,rw", CASES)
def test_device_sessions_equal_the_reference_spawner_on_the_same_draws(ref, base, overrides, st, rw, monkeypatch):
    import yaml
    from ev2gym_b200.reference_export import scenario_from_env, spawn_tables_from_env, topology_from_env
    cfg = yaml.safe_load(open(f"{REF}/ev2gym/example_config_files/{base}.yaml"))
    cfg.update(overrides)
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    try:
        env = ref.EV2Gym(config_file=f.name, seed=7, state_function=getattr(ref.state, st), reward_function=getattr(ref.reward, rw))
    finally:
        os.unlink(f.name)
    # a bank of S scenarios (time series + start dates from the reference's own reset()); the sessions are then re-drawn
    S, starts, scns = 6, [], []
    for i in range(S):
        env.reset(seed=100 + i)
        starts.append(env.sim_date)
        scns.append(scenario_from_env(env))
    topo = topology_from_env(env)
    tab = spawn_tables_from_env(env, starts)
    T, P = topo.T, topo.P
    monkeypatch.setenv("EV2B_KERNEL", "evlist")
    eng = ref.emu.EmuEngine(topo, S, reward=rw, state=st, outputs=("reward",))
    eng.set_spawn_tables(tab)
    eng.load_scenarios(scns)
    seed = 0xC0FFEE123457
    eng.resample_sessions(seed=seed)
    cdf = np.cumsum(np.asarray(tab.model_prob, dtype=np.float64) / float(np.sum(tab.model_prob)))
    port_off = np.asarray(topo.cs_port_off)
    total = 0
    for s in range(S):
        ctx = types.SimpleNamespace(c=None, n_normal=0, required=None, n_rand=0, has_lut=False)

        def rand(*shape):
            if len(shape) == 2:                                    # arrival_probabilities[port, t]          utils.py:489-490
                assert shape == (P, T)
                return np.array([[_uniform(seed, ((s * P + pt) * T + t) * 16) for t in range(T)] for pt in range(P)])
            assert shape == () and ctx.c is not None
            # scalar efficiencies (only without an efficiency curve) come first, the transition SoC last   :293-296, 309-310
            order = [9] if ctx.has_lut else [10, 11, 9]
            k = order[ctx.n_rand]
            ctx.n_rand += 1
            return _uniform(seed, ctx.c + k)

        def normal(mean, sd):                                      # required energy, then time of stay     :207-208, 236-237
            k = (1, 7)[ctx.n_normal]
            ctx.n_normal += 1
            v = _normal(seed, ctx.c + k, float(mean), float(sd))
            if k == 1:
                ctx.required = v
            return v

        def randint(lo, hi):
            if (lo, hi) == (5, 10) and ctx.n_normal == 1 and ctx.required is not None and ctx.required < 5:   # :210-211
                ctx.required = _randint(seed, ctx.c + 3, 5, 10)
                return ctx.required
            k = 5 if (not ctx.first_cap_drawn and hi < ctx.required) else 6                                  # :220-226
            ctx.first_cap_drawn = True
            return _randint(seed, ctx.c + k, int(lo), int(hi))

        def choice(names, p=None):                                 # np.random.choice(models, p=registrations)  :213-216
            u = _uniform(seed, ctx.c + 4)
            m = 0
            while m < len(cdf) - 1 and u >= cdf[m]:
                m += 1
            ctx.has_lut = bool(tab.model_lut[m] >= 0)
            return names[m]

        original = ref.utils.spawn_single_EV

        def spawn_with_context(env, scenario, cs_id, port, hour, minute, step, min_time_of_stay_steps):
            flat = int(port_off[cs_id]) + int(port)                # EV_spawner's running port counter
            ctx.c = ((s * P + flat) * T + int(step)) * 16
            ctx.n_normal, ctx.n_rand, ctx.required, ctx.first_cap_drawn = 0, 0, None, False
            ctx.has_lut = bool(tab.model_lut[0] >= 0) if not tab.heterogeneous else False
            return original(env, scenario, cs_id, port, hour, minute, step, min_time_of_stay_steps)

        monkeypatch.setattr(np.random, "rand", rand)
        monkeypatch.setattr(np.random, "normal", normal)
        monkeypatch.setattr(np.random, "randint", randint)
        monkeypatch.setattr(np.random, "choice", choice)
        monkeypatch.setattr(ref.utils, "spawn_single_EV", spawn_with_context)
        d0 = starts[s]
        env.sim_date = datetime.datetime(d0.year, d0.month, d0.day, d0.hour, d0.minute)
        evs = ref.utils.EV_spawner(env)
        monkeypatch.undo()
        monkeypatch.setenv("EV2B_KERNEL", "evlist")
        d = eng.read_sessions(s)
        loc = np.searchsorted(port_off, d["port"], side="right") - 1
        got = sorted(zip(d["t_arr"].tolist(), loc.tolist(), d["t_dep"].tolist(), np.round(d["cap0"], 9).tolist(),
                         tab.model_B[d["model"]].tolist()))
        want = sorted((int(ev.time_of_arrival), int(ev.location), int(ev.time_of_departure),
                       round(float(ev.battery_capacity_at_arrival), 9), float(ev.battery_capacity)) for ev in evs)
        assert len(got) == len(want), (s, len(got), len(want))
        for g, w in zip(got, want):
            assert g[:3] == w[:3] and abs(g[3] - w[3]) <= 1e-9 and g[4] == w[4], (s, g, w)
        if tab.heterogeneous:                                      # transition SoC / scalar efficiencies of every EV
            ts_got = sorted(np.round(d["ts"], 3).tolist())
            ts_want = sorted(round(float(ev.transition_soc), 3) for ev in evs)
            assert ts_got == ts_want, s
        total += len(want)
    assert total > 100, total
    eng.close()
