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


def bind_to_gpu_numa(local_rank: int) -> Optional[str]:
    """Pin this process to the CPUs that are NUMA-local to GPU ``local_rank`` (from the PCI device's
    ``local_cpulist`` in sysfs) so that pinned host buffers allocated afterwards are first-touched on the
    GPU's socket and H2D copies do not cross the inter-socket link.  Returns the cpulist applied, or None when
    the topology is not exposed (containers without sysfs PCI entries) - never raises."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id  # torch >= 2.6
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception:
        return None


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


# ---------------------------------------------------------------------------------------------------
# data-parallel training (SURVEY.md 8e "Training"): same batch sharding, ONE gradient all-reduce per step
# ---------------------------------------------------------------------------------------------------
def _real_view(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


class GradientAllReducer:
    """Average parameter gradients over the ranks after ``loss.backward()`` (train.py:329), before
    ``optimizer.step()`` (train.py:333) - what DDP would do to the reference's single-GPU step.

    Gradients (complex64 spectral weights viewed as fp32 pairs) are packed into flat fp32 buckets, each bucket
    is all-reduced asynchronously (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests) and scattered back
    divided by the world size.  With NVSwitch the cost is launch-latency bound, not link bound, so the default
    bucket is large (256 MB: the whole fsi FNO-2D model is one 268 MB message).
    BatchNorm batch statistics stay per rank (the reference has no SyncBN); ``sync_buffers`` broadcasts rank 0's
    running statistics like DDP's buffer broadcast."""

    def __init__(self, module: torch.nn.Module, dist, bucket_bytes: int = 256 << 20):
        self.module, self.dist = module, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.buckets, cur, cur_n = [], [], 0
        for p in self.params:
            n = _real_view(p).numel()
            if cur and (cur_n + n) * 4 > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += n
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)

    def sync_parameters(self, src: int = 0) -> None:
        """Broadcast rank ``src``'s parameters and buffers (start of training / after loading a checkpoint)."""
        if self.dist is None:
            return
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            self.dist.broadcast(_real_view(t.data), src=src)

    def sync_buffers(self, src: int = 0) -> None:
        if self.dist is None:
            return
        for t in self.module.buffers():
            self.dist.broadcast(t.data, src=src)

    @torch.no_grad()
    def __call__(self) -> int:
        """All-reduce + average every ``.grad``.  Returns the number of bytes reduced (0 on a single rank)."""
        if self.dist is None or self.world == 1:
            return 0
        works, total = [], 0
        for i, bucket in enumerate(self.buckets):
            grads = [_real_view(p.grad).reshape(-1) for p in bucket]
            n = sum(g.numel() for g in grads)
            flat = self._flat[i]
            if flat is None or flat.numel() != n or flat.device != grads[0].device:
                flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=grads[0].device)
            torch.cat(grads, out=flat)
            works.append(self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, async_op=True))
            total += n * 4
        inv = 1.0 / self.world
        for i, bucket in enumerate(self.buckets):
            works[i].wait()
            off = 0
            for p in bucket:
                g = _real_view(p.grad)
                g.copy_(self._flat[i][off:off + g.numel()].view_as(g)).mul_(inv)
                off += g.numel()
        return total
