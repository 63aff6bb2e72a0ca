"""world_size-2 gloo test (CPU) of the stream-sharding plumbing used by bench.py --gpus N (no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rstnet_b200.dist import reduce_timing, shard_streams


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_streams(2048, rank, world)
    ms, frames = reduce_timing(10.0 + 5.0 * rank, len(mine) * 125)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, ms, frames, sorted(sum(gathered, []))))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, frames, allstreams in res:
        assert ms == 15.0                       # max over ranks
        assert frames == 2048 * 125             # whole-job units
        assert allstreams == list(range(2048))  # every stream owned exactly once


def test_shard_streams_balanced():
    for world in (1, 2, 4, 8):
        sizes = [len(shard_streams(2048, r, world)) for r in range(world)]
        assert sum(sizes) == 2048 and max(sizes) - min(sizes) <= 1
