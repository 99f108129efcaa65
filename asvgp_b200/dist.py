"""Data-parallel plumbing: N shards over ranks (one process per GPU), ONE sum-allreduce of the packed accumulator
[G band | b | sum y^2 | count] over NCCL/NVLink, factorisation replicated per rank.  Prediction shards
the test points and needs no collective.  Works with any torch.distributed backend (the CPU tests use gloo)."""
import torch
import torch.distributed as dist


def is_distributed(flag="auto"):
    if flag is False:
        return False
    ok = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if flag is True and not ok:
        raise RuntimeError("distributed=True but torch.distributed is not initialised with world_size > 1")
    return ok


def shard_bounds(n, rank, world, align=2):
    """Contiguous [lo, hi) slice of n points for `rank`; boundaries are multiples of `align` points so that every
    shard of a 16-byte aligned array stays 16-byte aligned (the accumulate kernel's 128-bit load path)."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def allreduce_packed(acc, group=None):
    """In-place SUM all-reduce of the packed accumulator; returns it."""
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
