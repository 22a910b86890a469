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


def _parse_cpulist(cpulist: str) -> set:
    cpus = set()
    for part in cpulist.strip().split(","):
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def _pci_address(index: int) -> Optional[str]:
    """'dddd:bb:dd.f' of CUDA device ``index`` (torch device properties, else nvidia-smi)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        if hasattr(pr, "pci_bus_id"):
            return f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{getattr(pr, 'pci_device_id', 0):02x}.0"
    except Exception:
        pass
    try:
        import subprocess
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        r = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True,
                           text=True, timeout=20)
        rows = [l.split(",") for l in r.stdout.strip().splitlines()]
        ids = {int(a): b.strip().lower() for a, b in rows}
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        addr = ids[phys]  # nvidia-smi prints an 8-digit domain
        dom, rest = addr.split(":", 1)
        return f"{int(dom, 16):04x}:{rest}"
    except Exception:
        return None


def gpu_host_locality(index: int) -> dict:
    """PCI address, NUMA node and NUMA-local CPUs of CUDA device ``index`` as far as sysfs exposes them."""
    info = {"gpu": index, "pci": _pci_address(index), "numa_node": None, "local_cpulist": None}
    if info["pci"]:
        base = f"/sys/bus/pci/devices/{info['pci']}"
        try:
            with open(base + "/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
        except (OSError, ValueError):
            pass
        try:
            with open(base + "/local_cpulist") as f:
                info["local_cpulist"] = f.read().strip()
        except OSError:
            pass
    if info["numa_node"] is not None and info["numa_node"] >= 0 and not info["local_cpulist"]:
        try:
            with open(f"/sys/devices/system/node/node{info['numa_node']}/cpulist") as f:
                info["local_cpulist"] = f.read().strip()
        except OSError:
            pass
    return info


_MPOL_DEFAULT, _MPOL_PREFERRED, _MPOL_BIND = 0, 1, 2


def set_memory_policy(node: Optional[int], strict: bool = False) -> bool:
    """``set_mempolicy(2)`` for the calling thread: pages touched / pinned afterwards come from NUMA node ``node``
    (MPOL_PREFERRED, or MPOL_BIND when ``strict``); ``None`` restores the default first-touch policy.  This works when
    the container's cpuset confines the CPUs to one socket but not the memory - the case where CPU affinity alone
    cannot place a pinned buffer next to the GPU.  Returns False when the kernel refuses (never raises)."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        SYS_set_mempolicy = 238  # x86_64
        if os.uname().machine == "aarch64":
            SYS_set_mempolicy = 237
        if node is None or node < 0:
            return libc.syscall(SYS_set_mempolicy, _MPOL_DEFAULT, None, 0) == 0
        nbits = 1024
        mask = (ctypes.c_ulong * (nbits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
        w = 8 * ctypes.sizeof(ctypes.c_ulong)
        mask[node // w] |= 1 << (node % w)
        mode = _MPOL_BIND if strict else _MPOL_PREFERRED
        return libc.syscall(SYS_set_mempolicy, mode, mask, nbits + 1) == 0
    except Exception:
        return False


def bind_to_gpu_numa(local_rank: int) -> Optional[str]:
    """Place this process next to GPU ``local_rank``: (1) restrict its CPUs to the GPU's NUMA-local ones when they
    are a proper subset of what the process may use, (2) prefer the GPU's NUMA node for memory allocated from now
    on (the pinned staging buffers), so H2D copies do not cross the inter-socket link.  Returns a description of what
    was applied, or None when the topology is not exposed - never raises."""
    try:
        info = gpu_host_locality(local_rank)
        applied = []
        if info["local_cpulist"]:
            allowed = os.sched_getaffinity(0)
            cpus = _parse_cpulist(info["local_cpulist"]) & allowed
            if cpus and cpus != allowed:
                os.sched_setaffinity(0, cpus)
                applied.append(f"cpus {info['local_cpulist']}")
        node = info["numa_node"]
        if node is not None and node >= 0 and set_memory_policy(node):
            applied.append(f"mempolicy preferred node{node}")
        return ", ".join(applied) if applied else None
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


def _broadcast_into(module, tensors, dist, src: int) -> None:
    """Broadcast ``tensors`` of ``module`` from rank ``src`` in place.  The writes go through ``.data`` (leaf
    parameters that require grad cannot be written in place otherwise), which does not advance the tensors' version
    counters - so the engine's packed copy (keyed by data_ptr / version / ``_stats_epoch``) is invalidated
    explicitly: the next forward on every rank re-packs from the synchronised values."""
    if dist is None:
        return
    for t in tensors:
        dist.broadcast(_real_view(t.data), src=src)
    module._stats_epoch = getattr(module, "_stats_epoch", 0) + 1


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
        _broadcast_into(self.module, list(self.module.parameters()) + list(self.module.buffers()), self.dist, src)

    def sync_buffers(self, src: int = 0) -> None:
        """DDP's buffer broadcast: rank ``src``'s BatchNorm running statistics to every rank."""
        _broadcast_into(self.module, list(self.module.buffers()), self.dist, src)

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


class OverlappedGradientReducer:
    """Gradient averaging overlapped with the backward pass (GPU / NCCL only).

    ``b200fno_train_backward`` records a CUDA event as soon as each group of gradients is final (projection first,
    then the Fourier layers from last to first).  This object is attached to the module; the engine's autograd
    function calls :meth:`backward_and_reduce`, which enqueues the whole backward on the compute stream, then issues
    the all-reduce of every group on a side stream gated by that group's event - so NCCL moves layer ``l``'s 67-134 MB
    of spectral-weight gradients over NVLink while the kernels of layers ``l-1 .. 0`` are still running - and finally
    makes the compute stream wait for the reductions.  The gradients autograd hands to ``.grad`` are already averaged:
    ``train.py:329-333`` needs no extra call (``ReduceOp.AVG``, no flatten copies for the large tensors)."""

    def __init__(self, module: torch.nn.Module, dist):
        self.module, self.dist = module, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.n_layers = module.n_layers
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("OverlappedGradientReducer needs the module on a CUDA device (use GradientAllReducer)")
        self.side = torch.cuda.Stream(device=dev)
        # The reductions share the SMs with the backward pass.  A stock NCCL communicator launches up to 32 CTAs per
        # collective on NVLink systems, and while such a kernel waits for a slower peer it holds those SMs: measured on
        # 8 B200 (profiles/r02_train_bench_n8.json) the overlapped step took 19.2 ms against 11.1 ms with the
        # all-reduce after the backward.  The overlapped collectives therefore run on their OWN communicator capped at
        # a few CTAs (NVSwitch needs little: 537 MB per step under a 5 ms backward is ~110 GB/s) - B200FNO_REDUCER_MAX_CTAS.
        self.group = None
        max_ctas = int(os.environ.get("B200FNO_REDUCER_MAX_CTAS", "8"))
        if dist is not None and self.world > 1 and max_ctas > 0:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = max_ctas
                opts.config.min_ctas = min(4, max_ctas)
                self.group = dist.new_group(backend="nccl", pg_options=opts)
            except Exception:  # an older torch without ncclConfig: fall back to the default communicator
                self.group = None
        self.max_ctas = max_ctas if self.group is not None else 0
        self.events = [torch.cuda.Event() for _ in range(self.n_layers + 1)]
        for e in self.events:  # torch creates the underlying cudaEvent_t lazily on the first record
            e.record(torch.cuda.current_stream(dev))
        self.bytes_last = 0
        module._grad_sync = self if self.world > 1 else None

    def detach(self) -> None:
        self.module._grad_sync = None

    def sync_parameters(self, src: int = 0) -> None:
        _broadcast_into(self.module, list(self.module.parameters()) + list(self.module.buffers()), self.dist, src)

    def sync_buffers(self, src: int = 0) -> None:
        _broadcast_into(self.module, list(self.module.buffers()), self.dist, src)

    def _group(self, names, idx):
        L = self.n_layers
        if idx == L:
            return [n for n in names if n.startswith(("fc1.", "fc2."))]
        tag = f".{idx}."
        return [n for n in names if tag in n and n.startswith(("spectral_convs.", "convs.", "bns."))]

    def backward_and_reduce(self, engine, x, dy, params: dict, seq=None, input_grad: bool = False) -> dict:
        dist, L = self.dist, self.n_layers
        cur = torch.cuda.current_stream(x.device)
        handles = [e.cuda_event for e in self.events]
        if not all(handles):
            raise RuntimeError("OverlappedGradientReducer: a gradient-ready event has no cudaEvent_t handle")
        grads = engine.train_backward(x, dy, params, ready_events=handles, seq=seq, input_grad=input_grad)
        dx = grads.pop("__input__", None)  # the input gradient is per shard: not part of any reduction group
        if os.environ.get("B200FNO_REDUCER_SYNC"):  # race probe (profiles/diag_train.py): no overlap at all
            torch.cuda.synchronize(x.device)
        if os.environ.get("B200FNO_REDUCER_SNAPSHOT"):  # diag: the local gradients as the backward left them
            self.debug_snapshot = {n: g.clone() for n, g in grads.items()}
        done = torch.cuda.Event()
        done.record(cur)  # fc0 gradients (and everything else) final
        works, total = [], 0
        op = dist.ReduceOp.AVG
        dbg = os.environ.get("B200FNO_REDUCER_MODE", "")  # diag only: "nonccl" skips the collectives

        class _NoWork:
            def wait(self):
                pass

        def all_reduce(t):
            return _NoWork() if dbg == "nonccl" else dist.all_reduce(t, op=op, group=self.group, async_op=True)

        with torch.cuda.stream(self.side):
            for idx in [L] + list(range(L - 1, -1, -1)) + [-1]:
                self.side.wait_event(self.events[idx] if idx >= 0 else done)
                names = self._group(grads.keys(), idx) if idx >= 0 else [n for n in grads if n.startswith("fc0.")]
                small = [grads[n] for n in names if grads[n].numel() < (1 << 16)]
                for n in names:
                    g = grads[n]
                    if g.numel() >= (1 << 16):  # spectral weights: reduced in place, no flatten copy
                        if dbg == "smallonly":
                            continue
                        if dbg == "dummy":  # same NCCL traffic on memory nobody else uses
                            if not hasattr(self, "_dummy"):
                                self._dummy = {}
                            t = self._dummy.setdefault(n, torch.zeros_like(_real_view(g)))
                            works.append((all_reduce(t), None, None))
                        elif dbg == "clone":
                            c = _real_view(g).clone()
                            works.append((all_reduce(c), c, [g]))
                        else:
                            works.append((all_reduce(_real_view(g)), None, None))
                        total += _real_view(g).numel() * 4
                if small:  # biases, BatchNorm affine, 1x1 conv: one coalesced message per group
                    flat = torch.cat([_real_view(g).reshape(-1) for g in small])
                    works.append((all_reduce(flat), flat, small))
                    total += flat.numel() * 4
            for w, flat, small in works:
                w.wait()  # the side stream waits for NCCL
                if flat is not None:
                    off = 0
                    for g in small:
                        v = _real_view(g)
                        v.copy_(flat[off:off + v.numel()].view_as(v))
                        off += v.numel()
            fin = torch.cuda.Event()
            fin.record(self.side)
        cur.wait_event(fin)  # gradients returned to autograd are averaged as far as stream order goes
        for g in grads.values():
            g.record_stream(self.side)
        self.bytes_last = total
        if dx is not None:
            grads["__input__"] = dx
        return grads
