"""CPU test of the N > 1 host logic with torch.distributed (gloo, world_size 2): frame sharding covers the clip
exactly once with no overlap, and the max-over-ranks timing reduction works.  No data-path collective exists."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, n_frames: int, q) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from vr180_convert_b200.shard import gather_shards, max_over_ranks, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_range(n_frames, world, rank)
    shards = gather_shards(mine, dist)
    slowest = max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, list(mine), [list(s) for s in shards], slowest))


@pytest.mark.parametrize("n_frames", [1024, 7])
def test_two_rank_sharding(n_frames):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    covered = results[0][1] + results[1][1]
    assert covered == list(range(n_frames))
    for _, _, shards, slowest in results:
        assert [i for s in shards for i in s] == list(range(n_frames))
        assert slowest == 11.0
