"""Data-parallel sharding of independent utterances over the GPUs of one box.

Every utterance (voice prefix + text + its own KV pages, Mimi state and noise stream) is independent
(the reference copies the voice state per call, `pocket_tts_mlx/models/tts_model.py:372-373`), so the only
multi-GPU strategy is replicas: one process per GPU, a full weight copy each, no collective on the data
path.  These helpers are the whole "parallel layer": a balanced static partition and, for drivers that want
one result list, an order-preserving gather over `torch.distributed` (NCCL on GPUs, gloo in CPU tests).
torch is imported lazily: the single-GPU runtime stays torch-free.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


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


def dist_env() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) as torchrun exports them; (1, 0, 0) outside a launcher."""
    import os
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def generate_sharded(model, model_states: Sequence, token_ids: Sequence[Sequence[int]], *, rank: int, world_size: int,
                     slots: int = 256, noise: Optional[Sequence] = None, seed: int = 0, gather: bool = False,
                     cost: Optional[Sequence[int]] = None, **kwargs):
    """The multi-GPU product path: this rank's share of an utterance set through the continuous-batching scheduler.

    One process per GPU, each holding a full replica (`model` lives on this rank's GPU).  Every rank computes the
    same deterministic partition (`shard_by_length` over the utterances' frame budgets, longest first, so replicas
    finish together), decodes only its own utterances with `model.generate_audio_continuous`, and no byte of the data
    path crosses GPUs.  Returns `(indices, waves)` -- the global indices this rank decoded and their waveforms -- or,
    with `gather=True`, the full list in input order on every rank (`gather_in_order`; meant for result lists that fit
    in host memory, not for benchmarks).  Per-utterance `noise` arrays make an utterance's result independent of the
    rank and slot it lands in; with only a `seed` every rank draws its own block noise from `seed + rank`.
    The reference decodes one utterance at a time on one device (`models/tts_model.py:346-361`)."""
    n = len(model_states)
    if len(token_ids) != n:
        raise ValueError("model_states and token_ids differ in length")
    if cost is None:
        cost = [model._estimate_max_gen_len(len(t)) for t in token_ids]
    mine = shard_by_length(list(cost), world_size)[rank]
    fae = kwargs.pop("frames_after_eos", 3)
    fae_mine = fae if isinstance(fae, int) else [fae[i] for i in mine]
    out = model.generate_audio_continuous(
        [model_states[i] for i in mine], [token_ids[i] for i in mine], slots=slots, frames_after_eos=fae_mine,
        noise=None if noise is None else [noise[i] for i in mine], seed=seed + rank, **kwargs)
    if not gather:
        return mine, out
    if isinstance(out, tuple):                      # return_latents=True: gather both lists
        return tuple(gather_in_order(list(part), mine, n) for part in out)
    return gather_in_order(list(out), mine, n)


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
