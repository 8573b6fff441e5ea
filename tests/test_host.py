"""CPU: host-side logic -- scenario containers, port assignment, the C-ABI library's exports, env sharding
and the KPI all-reduce on a world_size-2 gloo group.  No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from ev2gym_b200 import _lib
from ev2gym_b200.distributed import shard_range, shard_scenario_ids
from ev2gym_b200.scenario import ScenarioPack, Topology, assign_ports
from ev2gym_b200.synthetic import sample_bank


def test_pack_roundtrip(tmp_path):
    topo = Topology.uniform(C=6, n_ports=2, Tr=2, T=48)
    bank = sample_bank(topo, 3, seed=4, min_stay=4)
    p = str(tmp_path / "x.npz")
    ScenarioPack(topo, bank, "t").save(p)
    back = ScenarioPack.load(p)
    assert back.topo.P == topo.P and len(back) == 3
    for a, b in zip(bank, back.scenarios):
        assert np.array_equal(a.tr_infl, b.tr_infl) and np.array_equal(a.luts_c, b.luts_c)
        for k in a.sessions:
            assert np.array_equal(a.sessions[k], b.sessions[k], equal_nan=True), k


def test_assign_ports_first_free_rule():
    topo = Topology.uniform(C=1, n_ports=2, Tr=1, T=30)
    # EV0 (3..10) takes port 0, EV1 (4..6) port 1, EV2 arrives at 7: port 1 is free again (EV1 left in step 6)
    port = assign_ports(topo, np.array([3, 4, 7, 11]), np.array([10, 6, 20, 15]), np.zeros(4, int))
    assert list(port) == [0, 1, 1, 0]          # EV3 arrives at 11: EV0 left during step 10 -> first free is port 0
    with pytest.raises(ValueError):
        assign_ports(topo, np.array([3, 4, 5]), np.array([10, 10, 10]), np.zeros(3, int))
    with pytest.raises(ValueError):
        assign_ports(topo, np.array([5, 3]), np.array([10, 10]), np.zeros(2, int))


def test_library_exports_every_declared_symbol():
    """include/ev2b.h <-> libev2b.so: every declared entry point is exported (built by __graft_entry__.build)."""
    if _lib.needs_build():
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "ev2b.h")).read()
    declared = set(re.findall(r"\b(ev2b_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym
    L.ev2b_abi_version.restype = ctypes.c_int
    assert L.ev2b_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (no oracle / CPU route)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ev2gym_b200.engine import BatchedEngine, EngineError
    with pytest.raises(EngineError):
        BatchedEngine(Topology.uniform(C=2, n_ports=1, Tr=1, T=8), 1)
    src = "".join(open(os.path.join(ROOT, "ev2gym_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "ev2gym_b200"))
                  if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_shard_range_partitions_exactly():
    for total, world in [(4096, 8), (10, 3), (7, 8), (1, 1)]:
        parts = [shard_range(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        ids = sum((shard_scenario_ids(total, r, world, 5) for r in range(world)), [])
        assert ids == [g % 5 for g in range(total)]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from ev2gym_b200.distributed import allreduce_kpis, kpi_dict, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    local = torch.zeros(13, dtype=torch.float64)
    local[0] = float(sum(range(lo, hi)))            # stands in for the per-rank KPI sums
    local[12] = hi - lo
    allreduce_kpis(local)
    q.put((rank, kpi_dict(local)))
    dist.destroy_process_group()


def test_kpi_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    for _, k in res:
        assert k["total_reward"] == float(sum(range(10))) and k["steps"] == 10.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/ev2gym"), reason="needs the reference checkout (build container only)")
def test_replay_pickle_import_round_trip():
    """EvCityReplay import (ev2gym/models/replay.py): `scenario_from_replay` == exporting the env the reference builds
    from the same pickle, and the oracle on it reproduces the reference's re-run bit for bit.  Runs the reference in a
    subprocess (tools/make_golden.py --replay-check changes directory and imports the stubs)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "make_golden.py"), "--replay-check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rewards bit-equal" in r.stdout


def test_step_kernels_keep_their_register_budget():
    """step_kernel is register-limited to 8 CTAs of 128 threads per SM (64 registers per thread), the lean event-driven
    kernel to 7 (72 registers: measured faster than 64 and 80 on the B200, profiles/r2_ab_registers.jsonl), its HEAVY
    instantiation (statistics / grid / per-port outputs) to 4 (128); the kernel parameter block must fit the 4 KB
    parameter space.  Read from the built library with cuobjdump (no GPU needed), so an edit that bloats them fails
    here and not as a slowdown on the GPU box."""
    import re
    import shutil
    import subprocess
    from ev2gym_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _lib.load()
    txt = subprocess.check_output([cuobjdump, "-res-usage", _lib.LIB_PATH], text=True)
    res = dict(re.findall(r"Function (\S+):\s*\n\s*(REG:\d+ STACK:\d+[^\n]*)", txt))
    checked = 0
    for name, usage in res.items():
        lean_step = "step_kernelIfLi2ELb1ELi128ELi8ELb0ELb0" in name          # c3: float actions, 2 ports, no optional outputs
        evl = re.search(r"evl_step_kernelI[fd]Li\dELb[01]ELi\dELb([01])ELb[01]ELi\d+EEEvNS_6ParamsE", name)   # <ActT, NP, UNI, G, HEAVY, KSTEP, TPB>
        if not (lean_step or evl):
            continue
        reg, stack = int(re.search(r"REG:(\d+)", usage).group(1)), int(re.search(r"STACK:(\d+)", usage).group(1))
        const0 = int(re.search(r"CONSTANT\[0\]:(\d+)", usage).group(1))
        limit = 64 if lean_step else (128 if evl.group(1) == "1" else 72)
        assert reg <= limit and stack <= 256, (name, usage)
        assert const0 <= 4096 + 528, (name, usage)                            # driver area + kernel parameters
        checked += 1
    assert checked >= 4
