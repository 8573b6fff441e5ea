#!/usr/bin/env python
"""Differential fuzzing of the two step kernels on the SIMT emulator (tests/simt_emu): random topologies (1..3 ports or
ragged chargers, 1..6 transformers), batch sizes, reward / state functions, group sizes G, action dtypes, output sets and
thread schedules; evl_step_kernel against the C oracle every step (battery levels and masks exact, reward 1e-9, obs 1e-5).

    python tools/fuzz_emu_kernels.py [--trials 200] [--seed 0]

Test infrastructure only (no GPU needed)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))

REWARDS = ["SquaredTrackingErrorReward", "ProfitMax_TrPenalty_UserIncentives", "profit_maximization",
           "SqTrError_TrPenalty_UserIncentives", "SimpleReward", "MinimizeTrackerSurplusWithChargeRewards",
           "V2G_profitmax", "V2G_costs_simple", "V2G_profitmaxV2", "pst_V2G_profitmaxV2",
           "SquaredTrackingErrorRewardWithPenalty", None]
STATES = ["PublicPST", "V2G_profit_max", "V2G_profit_max_loads", None]


def close(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def trial(k, rng):
    import emu_engine
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch
    C = int(rng.integers(1, 70))
    n = int(rng.integers(1, 4))
    Tr = int(rng.integers(1, min(C, 6) + 1))
    T = int(rng.integers(8, 40))
    E = int(rng.integers(1, 10))
    G = int(rng.choice([1, 2, 4]))
    reward, state = REWARDS[rng.integers(len(REWARDS))], STATES[rng.integers(len(STATES))]
    adt = rng.choice(["float32", "float64"])
    topo = Topology.uniform(C=C, n_ports=n, Tr=Tr, T=T, imin=float(rng.choice([0.0, 6.0])))
    ragged = rng.random() < 0.3
    if ragged:
        topo.cs_n_ports[:] = rng.integers(1, 5, C)
        topo.cs_imax[rng.random(C) < 0.5] = 16.0
    outs = ["reward", "status"] + [o for o in ("obs", "tr_power", "tr_overload", "cs_power", "cs_current", "total_costs",
                                               "action_mask") if rng.random() < 0.7]
    os.environ["EV2B_KERNEL"], os.environ["EV2B_EVL_G"] = "evlist", str(G)
    os.environ["EV2B_EVL_TPB"] = "32" if (G == 1 and rng.random() < 0.4) else "128"     # one env per CTA / four warps per CTA
    desc = dict(k=k, C=C, n=n, ragged=ragged, Tr=Tr, T=T, E=E, G=G, reward=reward, state=state, adt=str(adt), outs=outs,
                tpb=os.environ["EV2B_EVL_TPB"])
    bank = sample_bank(topo, 3, seed=int(rng.integers(1 << 30)), min_stay=int(rng.integers(1, 6)),
                       occupancy=float(rng.uniform(0.1, 0.9)))
    eng = emu_engine.EmuEngine(topo, E, reward=reward, state=state, outputs=tuple(outs))
    eng.load_scenarios(bank)
    ids = [int(x) for x in rng.integers(0, 3, E)]
    eng.reset(scn_ids=ids)
    orc = OracleBatch(topo, [bank[i] for i in ids], reward=reward, state=state)
    orc.reset()
    caps = eng.state()["port_cap"]
    for t in range(T + 1):                      # one step beyond the end: WAS_DONE
        a = rng.uniform(-1.3, 1.3, (E, topo.P))
        a[rng.random((E, topo.P)) < 0.15] = 0.0
        if rng.random() < 0.1:
            a[:] = 1.0
        a = np.ascontiguousarray(a.astype(adt))
        out = eng.step(a)
        if t == T:
            assert (out["status"] & 4).all(), (desc, "was_done")
            break
        orc.step(a.astype(np.float64))
        occ = orc.arr["port_session"] >= 0
        assert np.array_equal(caps[occ], orc.arr["port_cap"][occ]), (desc, t, "cap")
        assert close(out["reward"], orc.reward, 1e-9, 1e-9), (desc, t, "reward", out["reward"], orc.reward)
        if "obs" in out and eng.D:
            assert close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), (desc, t, "obs")
        if "action_mask" in out:
            assert np.array_equal(out["action_mask"] > 0, occ), (desc, t, "mask")
        if "tr_power" in out:
            assert close(out["tr_power"], orc.o["tr_power"][:, :Tr], 1e-9, 1e-9), (desc, t, "tr_power")
        if "cs_power" in out:
            assert close(out["cs_power"], orc.o["cs_power"], 1e-5, 1e-6), (desc, t, "cs_power")
        if "cs_current" in out:
            assert close(out["cs_current"], orc.o["cs_current"], 1e-5, 1e-6), (desc, t, "cs_current")
        assert np.array_equal((out["status"] & 1) > 0, orc.done > 0), (desc, t, "done")
        ovf = np.array([o.error == 1 for o in orc.outs])
        assert np.array_equal((out["status"] & 2) > 0, ovf), (desc, t, "amps overflow flag")
    assert eng.kernel_launches()[0] == 0, (desc, eng.kernel_launches())
    eng.close()
    return desc


def stateful_trial(k, rng):
    """A random sequence of API calls on one handle -- steps with changing output sets (some force step_kernel, so the
    connected-EV list is re-derived), fresh output buffers (full observation / mask rewrites), partial resets in the middle of
    an episode, device-side auto reset, ev2b_step_host (two env chunks) -- mirrored on the oracle env by env."""
    import ctypes as C
    import emu_engine
    from ev2gym_b200.scenario import Topology
    from ev2gym_b200.synthetic import sample_bank
    from oracle.oracle import OracleBatch, _p
    C_ = int(rng.integers(2, 40))
    n = int(rng.integers(1, 3))
    Tr = int(rng.integers(1, min(C_, 4) + 1))
    T = int(rng.integers(10, 26))
    S = int(rng.integers(1, 4))
    E = S * int(rng.integers(1, 4))                  # every scenario of the bank is in use from the start
    G = int(rng.choice([1, 2, 4]))
    reward = REWARDS[rng.integers(len(REWARDS) - 1)]
    state = STATES[rng.integers(len(STATES) - 1)]
    os.environ["EV2B_KERNEL"], os.environ["EV2B_EVL_G"] = "evlist", str(G)
    os.environ["EV2B_EVL_TPB"] = "32" if (G == 1 and rng.random() < 0.4) else "128"
    import math
    stride = E % S or 1                              # ev2b_reset_done: next scenario = (current + stride) mod S  (include/ev2b.h)
    while S > 1 and math.gcd(stride, S) != 1:
        stride += 1
    topo = Topology.uniform(C=C_, n_ports=n, Tr=Tr, T=T)
    bank = sample_bank(topo, S, seed=int(rng.integers(1 << 30)), min_stay=int(rng.integers(1, 6)),
                       occupancy=float(rng.uniform(0.2, 0.9)))
    base = ("reward", "status", "obs", "action_mask", "cs_power", "tr_power")
    eng = emu_engine.EmuEngine(topo, E, reward=reward, state=state, outputs=base)
    eng.load_scenarios(bank)
    ids = [e % S for e in range(E)]
    eng.reset(scn_ids=ids)
    orc = OracleBatch(topo, [bank[i] for i in ids], reward=reward, state=state)
    orc.reset()
    caps = eng.state()["port_cap"]
    desc = dict(k=k, C=C_, n=n, Tr=Tr, T=T, S=S, E=E, G=G, reward=reward, state=state)
    graveyard = []

    def oracle_reset(envs):
        for e in envs:
            orc.L.ev2o_reset(C.byref(orc._t.c), orc._scn_ptrs[e], C.byref(orc.states[e]), orc.state_kind, _p(orc.o["obs"][e]))
            orc.done[e] = 0

    def check(out, what, host=False):
        occ = orc.arr["port_session"] >= 0
        live = orc.done == 0
        assert np.array_equal(caps[occ], orc.arr["port_cap"][occ]), (desc, what, "cap")
        assert close(out["reward"], orc.reward, 1e-9, 1e-9), (desc, what, "reward")
        assert close(out["obs"], orc.o["obs"][:, :eng.D], 1e-5, 1e-5), (desc, what, "obs")
        assert np.array_equal((out["status"] & 1) > 0, orc.done > 0), (desc, what, "done")
        if not host:
            assert np.array_equal(out["action_mask"] > 0, occ), (desc, what, "mask")
            assert close(out["cs_power"], orc.o["cs_power"], 1e-5, 1e-6), (desc, what, "cs_power")
            assert close(out["tr_power"], orc.o["tr_power"][:, :Tr], 1e-9, 1e-9), (desc, what, "tr_power")
        return live

    for it in range(int(rng.integers(20, 70))):
        op = rng.choice(["step", "step", "step", "step_v6", "fresh", "partial_reset", "host"])
        if os.environ.get("FUZZ_TRACE"):
            print("  op", it, op, flush=True)
        a = rng.uniform(-1.1, 1.1, (E, topo.P))
        a[rng.random((E, topo.P)) < 0.15] = 0.0
        a = np.ascontiguousarray(a.astype(np.float32))
        if op == "fresh":
            graveyard.append(eng.out)                # keep the old buffers alive: the new ones must get NEW addresses (the
            eng.set_outputs(base)                    # library rewrites a buffer in full only when its pointer changes)
            for v in eng.out.values():
                v[...] = 1                           # stale content: obs / mask rows must be rewritten in full
            continue
        if op == "partial_reset" and E > 1:
            lo = int(rng.integers(0, E - 1))
            hi = int(rng.integers(lo + 1, E + 1))
            eng.reset(lo, hi, scn_ids=ids[lo:hi])
            oracle_reset(range(lo, hi))
            continue
        if op == "host":
            rew, st = np.zeros(E), np.zeros(E, dtype=np.uint32)
            obs = np.zeros((E, max(eng.D, 1)), dtype=np.float32)
            eng.step_host(a, rew, st, obs)
            orc.step(a.astype(np.float64))
            check({"reward": rew, "status": st, "obs": obs}, (it, op), host=True)
            graveyard.append(eng.out)
            eng.set_outputs(base)                    # (step_host advanced the envs without these buffers: start them afresh)
        else:
            if op == "step_v6":
                graveyard.append(eng.out)
                eng.set_outputs(base + ("port_energy",))
            out = eng.step(a)
            orc.step(a.astype(np.float64))
            check(out, (it, op))
            if op == "step_v6":
                graveyard.append(eng.out)
                eng.set_outputs(base)
        if orc.done.any():                           # finished envs restart on their NEXT scenario, on both sides
            eng.reset_done()
            fin = np.nonzero(orc.done)[0]
            for e in fin:
                ids[e] = (ids[e] + stride) % S
                orc._scn_ptrs[e] = C.pointer(orc._scn[ids[e]].c)      # (env i < S was built on scenario i)
            assert np.array_equal(eng.state()["env_scn"], ids), (desc, it, "scenario ids after the auto reset")
            oracle_reset(fin)
    eng.close()
    return desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stateful", action="store_true", help="random API call sequences instead of plain episodes")
    args = ap.parse_args()
    os.environ.setdefault("SIMT_EMU_SEED", str(1 + args.seed))
    rng = np.random.default_rng(args.seed)
    for k in range(args.trials):
        d = (stateful_trial if args.stateful else trial)(k, rng)
        if k % 20 == 0:
            print("ok", d, flush=True)
    print(f"{args.trials} trials passed")


if __name__ == "__main__":
    main()
