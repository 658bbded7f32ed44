"""Sharding of a batch of conditions over ranks and the single objective all-reduce.

The path is embarrassingly parallel over conditions (the reference's only parallelism is the OpenMP loop over
``edatas``, reference src/amici.cpp:235-251): every rank integrates a contiguous slice and only
(sum llh, sum sllh[nplist], n_failed) is exchanged — one all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of ``n`` conditions for ``rank``; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_objective(obj):
    """Sum the (1 + nplist + 1)-vector (sum llh, sum sllh, n_failed) over all ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(obj, op=dist.ReduceOp.SUM)
    return obj
