"""Multi-GPU plumbing: env replicas are independent (no cross-env term anywhere in the step,
SURVEY.md section 8e), so the env index range is partitioned over ranks and nothing on the data
path crosses GPUs.  The ONLY collective is a sum all-reduce of the small float64 KPI vector
(total reward, profits, energy, overload, EVs served ...) at reporting time -- NCCL on GPUs,
gloo in the CPU tests.  One process per GPU (torchrun)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .engine import KPI_NAMES


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the global env index range owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_scenario_ids(total_envs: int, rank: int, world: int, bank_size: int) -> List[int]:
    """Scenario of global env g is g mod bank_size, whatever the number of ranks."""
    lo, hi = shard_range(total_envs, rank, world)
    return [g % bank_size for g in range(lo, hi)]


def allreduce_kpis(local_kpi_sums, group=None):
    """Sum a [K] float64 tensor of per-rank KPI sums over all ranks (in place) and return it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_kpi_sums, op=dist.ReduceOp.SUM, group=group)
    return local_kpi_sums


def kpi_dict(kpi_sums) -> Dict[str, float]:
    v = kpi_sums.detach().cpu().numpy() if hasattr(kpi_sums, "detach") else np.asarray(kpi_sums)
    return {n: float(v[i]) for i, n in enumerate(KPI_NAMES)}
