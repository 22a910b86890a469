"""One-process-per-GPU plumbing for the batch-sharded rollout.

Inference shards the batch axis: rank r owns ``per_gpu`` contiguous samples of the global batch, weights are
replicated and there is NO data-path collective.  The only communication is the reduction of the timing
scalar (max over ranks) and a barrier around the timed region; NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """Initialise torch.distributed from the torchrun environment; returns the module or None (world 1)."""
    rank, world, _ = env_rank_world()
    if world <= 1:
        return None
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the global batch owned by ``rank`` (remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, dist, device="cpu") -> float:
    if dist is None:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist, device="cpu") -> float:
    if dist is None:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, ms_this_rank: float, dist, device="cpu") -> Tuple[float, float]:
    """Whole-job rate: units of all ranks / max-over-ranks time.  Returns (units_per_second, max_ms)."""
    ms = max_over_ranks(ms_this_rank, dist, device)
    units = sum_over_ranks(units_this_rank, dist, device)
    return units / (ms * 1e-3), ms
