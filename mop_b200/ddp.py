"""Batch-sharded data parallelism for the MoP models (one process per GPU).

The attention path shards by batch: every (sample, head) problem is independent, so the
only collective is the gradient all-reduce of torch DDP over NCCL (NVLink/NVSwitch).  The
reference has no multi-GPU code; this module is the whole of ours.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group when WORLD_SIZE > 1.  nccl on GPU, gloo on CPU."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def wrap(model: torch.nn.Module, local_rank: int = 0, bucket_cap_mb: int = 64) -> torch.nn.Module:
    """DDP wrapper with a single large bucket for small models (4 M params = 16 MB: one all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    on_gpu = next(model.parameters()).is_cuda
    return torch.nn.parallel.DistributedDataParallel(
        model, device_ids=[local_rank] if on_gpu else None, bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)


def max_over_ranks(value: float, device) -> float:
    """Timing rule: report the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
