#!/usr/bin/env python
"""Golden sessions / setpoints of the DEVICE sampler, produced by the SIMT-emulated build of the CUDA sources (CPU):

    python tools/make_sampler_golden.py        # -> tests/golden/sampler_<bank>.npz

The emulated kernels are pinned to the reference by tests/test_spawner_reference.py and tests/test_setpoints_reference.py
(the unmodified EV_spawner / generate_power_setpoints fed the same random numbers return the same sessions / setpoints);
tests/test_spawn.py::test_gpu_sampler_equals_the_golden_sessions then checks that the real library on the GPU draws exactly
these sessions for the same seed (the only difference between the two builds is libm: log / cos / sqrt of Box-Muller)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
SEED = 0x5EEDBA5E2024
BANKS = (("c2_publicpst_c25", "SquaredTrackingErrorReward", "PublicPST", 12), ("c3_v2gloads_c100n2tr5", "ProfitMax_TrPenalty_UserIncentives", "V2G_profit_max_loads", 4))


def main():
    import emu_engine
    from ev2gym_b200.scenario import ScenarioPack, SpawnTables
    emu_engine.build()
    os.environ["EV2B_KERNEL"] = "evlist"
    data = os.path.join(ROOT, "ev2gym_b200", "data")
    for name, reward, state, S in BANKS:
        pack = ScenarioPack.load(os.path.join(data, name + ".npz"))
        tab = SpawnTables.load(os.path.join(data, "spawn_" + name + ".npz"))
        eng = emu_engine.EmuEngine(pack.topo, S, reward=reward, state=state, outputs=("reward",))
        eng.set_spawn_tables(tab)
        eng.load_scenarios(pack.scenarios[:S])
        eng.resample_sessions(seed=SEED)
        out = {"seed": np.array([SEED], dtype=np.uint64), "n_scenarios": np.array([S])}
        for s in range(S):
            d = eng.read_sessions(s)
            for k, v in d.items():
                out[f"s{s}_{k}"] = v
            out[f"s{s}_setpoint"] = eng.read_setpoints(s)
        eng.close()
        path = os.path.join(ROOT, "tests", "golden", "sampler_" + name + ".npz")
        np.savez_compressed(path, **out)
        print(path, sum(len(out[f"s{s}_port"]) for s in range(S)), "sessions")


if __name__ == "__main__":
    main()
