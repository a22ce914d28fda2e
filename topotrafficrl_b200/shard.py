"""Multi-GPU sharding of env instances: contiguous blocks of GLOBAL env ids, one process
per GPU, no collective on the step path.  The only exchange is the sum of the episode-statistics vector at logging
cadence (maps to ``Evaluation.after_all_episodes``, reference trainer/evaluation.py:325-333), through
``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

STAT_FIELDS = ("episodes", "total_return", "total_length", "crashes", "arrivals", "total_speed", "vehicle_steps", "env_steps",
               "spawn_capacity_rejects", "sync_resets")


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[first, last) global env ids owned by ``rank``; the first ``total_envs % world_size`` ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def all_reduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """Sum the per-rank episode statistics over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    vec = torch.tensor([float(stats[k]) for k in STAT_FIELDS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_FIELDS, (float(x) for x in vec.tolist())))
