"""Data-parallel plumbing for the stream-sharded hot path: streams are independent, each rank
holds a full weight replica and its own KV / carry state, and there is NO collective on the data path.
torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for the start barrier and for reducing the
per-rank timings (max over ranks) and counters."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_streams(n_streams: int, rank: int, world: int) -> List[int]:
    """Stream i lives on rank i mod world (round-robin keeps ranks balanced as streams come and go)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_streams, world))


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_timing(ms: float, frames: int, device=None) -> Tuple[float, int]:
    """(max over ranks of the device-timed region, sum over ranks of the frames processed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms, frames
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    n = torch.tensor([frames], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())
