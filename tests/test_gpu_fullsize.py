"""GPU, at BASELINE.json's full sizes (c2 1024 x 25, c3 4096 x 100 x 2, c4 8192 x 250): size-independent properties.

The scenario banks hold 64 reference-exported episodes; env e plays scenario e mod 64 and every env of a scenario gets
the same action sequence.  Then
  (A) replica consistency: all envs of one scenario must agree BITWISE in every output of every step, whatever CTA /
      SM / launch position they ran in (catches races, cross-env leakage, uninitialised shared memory);
  (B) the first 64 envs are checked against the oracle for the whole episode (battery levels bit-exact, reward 1e-9,
      observation 1e-5, done flags exact) -- with (A) that covers every env of the batch;
  (C) the KPI accumulators equal the sums of the per-step outputs.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("c2_publicpst_c25", 1024, "SquaredTrackingErrorReward", "PublicPST"),
         ("c3_v2gloads_c100n2tr5", 4096, "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads"),
         ("c4_v2gprofitmax_c250", 8192, "profit_maximization", "V2G_profit_max")]


@pytest.mark.parametrize("pack_name,E,reward,state", CASES)
def test_full_size_replicas_and_sampled_oracle(pack_name, E, reward, state):
    import torch
    from ev2gym_b200.engine import BatchedEngine
    from ev2gym_b200.scenario import ScenarioPack
    from oracle.oracle import OracleBatch
    pack = ScenarioPack.load(os.path.join(ROOT, "ev2gym_b200", "data", pack_name + ".npz"))
    topo, S = pack.topo, len(pack)
    assert E % S == 0
    eng = BatchedEngine(topo, E, reward=reward, state=state, outputs=("reward", "status", "obs", "action_mask"))
    eng.load_scenarios(pack.scenarios)
    obs0 = eng.reset()                                   # env e -> scenario e mod S
    orc = OracleBatch(topo, pack.scenarios, reward=reward, state=state)
    o0 = orc.reset()
    assert np.allclose(obs0[:S].cpu().numpy(), o0, rtol=1e-5, atol=1e-5)
    caps = eng.state_tensors()["port_cap"]
    low = -1.0 if topo.v2g_enabled else 0.0
    rng = np.random.default_rng(11)
    reward_sum = torch.zeros(E, dtype=torch.float64, device="cuda")

    def same_across_replicas(t):
        v = t.reshape((E // S, S) + tuple(t.shape[1:]))
        return bool((v == v[:1]).all())

    for step in range(topo.T):
        a = rng.uniform(low, 1.0, (S, topo.P))
        a[rng.random((S, topo.P)) < 0.1] = 0.0
        out = eng.step(torch.tensor(np.tile(a, (E // S, 1)), dtype=torch.float32, device="cuda"))
        reward_sum += out["reward"]
        for k in ("reward", "status", "obs", "action_mask"):                       # (A)
            assert same_across_replicas(out[k]), (step, k)
        assert same_across_replicas(caps), (step, "cap")
        orc.step(a.astype(np.float32).astype(np.float64))                          # (B) the engine saw fp32 actions
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(caps[:S].cpu().numpy()[occ], orc.arr["port_cap"][occ]), step
        r = out["reward"][:S].cpu().numpy()
        assert np.all(np.abs(r - orc.reward) <= 1e-9 + 1e-9 * np.abs(orc.reward)), step
        assert np.allclose(out["obs"][:S].cpu().numpy(), orc.o["obs"][:, :eng.D], rtol=1e-5, atol=1e-5), step
        assert np.array_equal((out["status"][:S].cpu().numpy() & 1).astype(bool), orc.done.astype(bool)), step
    assert bool((out["status"] & 1).all())
    k = eng.kpis()                                                                 # (C)
    assert np.allclose(k["total_reward"], reward_sum.cpu().numpy(), rtol=1e-9, atol=1e-9)
    assert np.allclose(k["total_reward"][:S], [s.total_reward for s in orc.states], rtol=1e-9, atol=1e-9)
