"""Static sharding of independent frames over the GPUs of one box.

The path has no exchange step: every stereo pair (every output pixel) is independent, so the only cross-rank
operations are the benchmark's barrier and the max-over-ranks of the device time.  No data-path collective.
"""
from __future__ import annotations


def shard_range(n_items: int, world_size: int, rank: int) -> range:
    """Contiguous block partition: rank r owns [r*ceil(n/ws), min(n, (r+1)*ceil(n/ws)))."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world_size / rank")
    per = -(-n_items // world_size)
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def max_over_ranks(value: float, dist=None, device="cpu") -> float:
    """Max of a per-rank scalar (device time of the timed region).  `dist` is torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_shards(rng: range, dist=None, device="cpu") -> list[range]:
    """Every rank's shard (used by tests / drivers to check that the partition covers the clip exactly once)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [rng]
    import torch

    mine = torch.tensor([rng.start, rng.stop], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [range(int(o[0]), int(o[1])) for o in out]
