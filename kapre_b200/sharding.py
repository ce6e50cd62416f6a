"""Batch sharding for multi-GPU runs (one process per GPU, ``torch.distributed``).

Every operation on the hot path is independent per batch item -- the decibel clamp takes its
maximum per item (kapre/backend.py:178-179) -- so ranks work on disjoint contiguous batch slices
with NO collective in the data path; the only collectives are the barrier / MAX-reduction a
benchmark needs for timing, and an optional result gather.
"""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the batch owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(value: float, device=None) -> float:
    """MAX over ranks of a Python float (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_shards(local, dim=0):
    """Optional result collection: all-gather equally sized shards along ``dim`` (NCCL for CUDA
    tensors, gloo for CPU tensors).  Not part of any timed region."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, dim=dim)
