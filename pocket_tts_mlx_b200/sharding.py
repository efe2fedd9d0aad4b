"""Data-parallel sharding of independent utterances over the GPUs of one box.

Every utterance (voice prefix + text + its own KV pages, Mimi state and noise stream) is independent
(the reference copies the voice state per call, `pocket_tts_mlx/models/tts_model.py:372-373`), so the only
multi-GPU strategy is replicas: one process per GPU, a full weight copy each, no collective on the data
path.  These helpers are the whole "parallel layer": a balanced static partition and, for drivers that want
one result list, an order-preserving gather over `torch.distributed` (NCCL on GPUs, gloo in CPU tests).
torch is imported lazily: the single-GPU runtime stays torch-free.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_length(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment so replicas finish together: returns per-rank utterance indices
    (each rank's list is sorted by decreasing length, so its lock-step waves are homogeneous)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += lengths[i]
    return out


def gather_in_order(local_items: list, local_indices: Sequence[int], n_total: int) -> list:
    """All ranks contribute (index, item) pairs; every rank gets the full list in original order."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = [None] * n_total
        for i, x in zip(local_indices, local_items):
            out[i] = x
        return out
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, list(zip(local_indices, local_items)))
    out = [None] * n_total
    for part in bucket:
        for i, x in part:
            out[i] = x
    return out


def max_over_ranks(value: float) -> float:
    """Timing convention of bench.py: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
