"""
Multi-GPU plumbing: one process per GPU (``torchrun``), batch-dimension sharding, optional final gather.

The hot path never communicates: element-wise arrays, NTT batches and Reed-Solomon codewords are
independent along their leading dimension, so each rank processes a contiguous shard
(SURVEY.md §8e).  The only collective is the optional gather of the result shards
(``all_gather_into_tensor`` over NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous, balanced shard owned by ``rank``; the first ``total % world`` ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def gather_rows(local, total_rows: int, group=None):
    """
    All-gather row shards ([rows_r, ...] tensors laid out by ``shard_bounds``) into the full [total_rows, ...]
    tensor on every rank.  Uneven shards are padded to the largest shard for the collective and trimmed after.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(total_rows, world)
    biggest = max(sizes)
    tail = tuple(local.shape[1:])
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest,) + tail, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    out = torch.empty((world * biggest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(s == biggest for s in sizes):
        return out
    parts = [out[r * biggest: r * biggest + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
