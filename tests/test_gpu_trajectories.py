"""GPU: the batched trajectory generator against episodes recorded from the reference with the same agent
(tools/make_golden.py stores obs0/obs, the in-place-masked actions, rewards and dones of every step -- exactly what
ev2gym/scripts/generate_trajectories.py:69-81 appends)."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = [("pst25_roundrobin_s5", "roundrobin"), ("pst12n2_roundrobin_s12", "roundrobin"), ("pst25_calap_s6", "calap"),
         ("loads_c10n2_calap_s13", "calap"), ("c1_afap_s42", "afap"), ("homog_ts1_afap_s4", "afap")]


@pytest.mark.parametrize("name,agent", CASES)
def test_trajectory_matches_reference_episode(name, agent):
    from ev2gym_b200.scenario import ScenarioPack
    from ev2gym_b200.trajectories import generate_trajectories
    pack = ScenarioPack.load(f"{GOLDEN}/{name}.scenario.npz")
    tr = np.load(f"{GOLDEN}/{name}.trace.npz")
    trajs = generate_trajectories(pack.topo, pack.scenarios, 5, agents=(agent,), reward=str(tr["reward_fn"]),
                                  state=str(tr["state_fn"]), max_envs=3)          # 2 batches: 3 + 2 envs
    assert len(trajs) == 5
    T = tr["reward"].shape[0]
    want_obs = np.concatenate([tr["obs0"][None], tr["obs"][:-1]])                 # state BEFORE each step
    for tj in (trajs[0], trajs[4]):
        assert tj["observations"].shape == (T, want_obs.shape[1]) and tj["actions"].shape == (T, pack.topo.P)
        assert np.array_equal(tj["actions"], tr["actions_eff"])                   # float64, bit-exact
        assert np.allclose(tj["observations"], want_obs, rtol=1e-5, atol=1e-5)    # obs are float32 on the device
        assert np.allclose(tj["rewards"], tr["reward"], rtol=1e-9, atol=1e-9)
        assert np.array_equal(tj["dones"], tr["done"])


def test_mixed_agents_alternate_like_the_reference_script():
    """Trajectory i: AFAP for even i, RoundRobin for odd i (generate_trajectories.py:63-66), scenario i mod bank."""
    from ev2gym_b200.scenario import ScenarioPack
    from ev2gym_b200.trajectories import generate_trajectories
    pack = ScenarioPack.load(f"{GOLDEN}/pst25_roundrobin_s5.scenario.npz")
    tr = np.load(f"{GOLDEN}/pst25_roundrobin_s5.trace.npz")
    trajs = generate_trajectories(pack.topo, pack.scenarios, 4, reward="SquaredTrackingErrorReward", state="PublicPST")
    for i in (1, 3):
        assert np.array_equal(trajs[i]["actions"], tr["actions_eff"])
    occ = np.concatenate([np.zeros((1, pack.topo.P)), tr["action_mask"][:-1]])     # ports occupied before each step
    for i in (0, 2):
        assert np.array_equal(trajs[i]["actions"], occ)                            # AFAP: ones, masked in place
