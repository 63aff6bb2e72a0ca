"""world_size-2 gloo test (CPU) of the stream-sharding plumbing used by bench.py --gpus N (no data-path collective)."""
import os
import socket
import subprocess
import sys

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


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) must print one JSON line with the
    metric / config of our arm, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e block."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "codec_frames_per_s" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["config"]["workload"].startswith("mimi_streaming_encode_decode") and line["config"]["streams_per_gpu"] == 256
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without CUDA the product arm must stop with a clear message instead of measuring anything."""
    if torch.cuda.is_available():
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0", "--no-lm"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode != 0 and "needs a GPU" in (r.stderr + r.stdout)
