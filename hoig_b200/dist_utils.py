"""One-process-per-GPU plumbing shared by bench.py and tests (torch.distributed, NCCL on GPUs, gloo on CPU).

The inference path shards by batch with no data-path collective (InstanceNorm is per sample): rank r owns the
contiguous sub-batch [r*B, (r+1)*B).  The only collectives are the timing barrier and a MAX-reduce of elapsed times.
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous sub-batch of rank `rank`; the global batch must divide evenly (512 -> 8 x 64)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not shard over {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_seed(base: int, rank: int) -> int:
    return base + rank


def reduce_max(values: Sequence[float], device) -> list:
    """MAX over ranks of a few scalars (elapsed times); identity when not distributed."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def throughput(images_per_rank_per_step: int, world: int, steps: int, max_ms: float) -> float:
    """Whole-job images/s: every rank processed its shard in at most `max_ms`."""
    return images_per_rank_per_step * world * steps / (max_ms / 1e3)
