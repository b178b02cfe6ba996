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


def bind_host_near_gpu(device_index: int) -> list[int] | None:
    """One process per GPU: run this process (and therefore first-touch its pinned staging buffers) on the CPUs of
    the NUMA node the GPU hangs off, read from sysfs (`/sys/bus/pci/devices/<bdf>/local_cpulist`).  With 8 ranks on a
    two-socket host the host<->device copies of `vr180_ctx_run` otherwise cross the socket interconnect for half of
    the GPUs.  Returns the CPU list it bound to, or None when the topology is not available (nothing is changed)."""
    import os

    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus: set[int] = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:  # noqa: BLE001 -- no sysfs / no such attribute / not permitted: leave the affinity alone
        return None
