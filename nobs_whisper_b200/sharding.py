"""Data-parallel partition of independent windows / utterances over one process per GPU
(SURVEY.md §8e).  There is no collective on the compute path: every rank transcribes its own
contiguous block of windows with its own engine replica; only the host gathers the segment
texts back in window order (BASELINE.json north_star: "no NCCL on the compute path and only the
host gathering segments in order")."""
from __future__ import annotations


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous block of item indices owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_in_order(local_items: list, n_items: int, group=None) -> list:
    """Every rank contributes the results of its shard_range; every rank receives the full list
    in item order.  Uses torch.distributed object gather on the host (gloo or nccl group)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        if len(local_items) != n_items:
            raise ValueError("single process must hold every item")
        return list(local_items)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(local_items) != len(shard_range(n_items, world, rank)):
        raise ValueError("local result count does not match this rank's shard")
    parts = [None] * world
    dist.all_gather_object(parts, list(local_items), group=group)
    out = []
    for r in range(world):
        if len(parts[r]) != len(shard_range(n_items, world, r)):
            raise RuntimeError(f"rank {r} returned {len(parts[r])} results for a shard of {len(shard_range(n_items, world, r))}")
        out.extend(parts[r])
    return out


def transcribe_sharded(engine, audios: list, language=None, vocabulary=None, beam_size: int = 0, group=None) -> list[str]:
    """Transcribe `audios` (the same list on every rank) data-parallel: this rank runs only its
    shard through `engine.transcribe_batch`, then the texts are gathered in order."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = shard_range(len(audios), world, rank)
    local = engine.transcribe_batch([audios[i] for i in mine], language=language, vocabulary=vocabulary, beam_size=beam_size) if len(mine) else []
    return gather_in_order(local, len(audios), group)
