"""Two-GPU (NCCL) test of data-parallel training: the all-reduce overlapped with the backward pass
(dist.OverlappedGradientReducer, gated by the events b200fno_train_backward records) must give the same
averaged gradients as the plain post-backward reducer, and both must equal the mean of the per-rank gradients.
Skipped on boxes with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import realpdebench_b200 as R
    from realpdebench_b200 import dist as D
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist = D.init("nccl", dev)
    ctor = (6, 8, 3, 32, (4, 24, 20, 3), (4, 24, 20, 3))
    torch.manual_seed(0)  # same initial weights on both ranks
    m = R.FNO2d(*ctor).to(dev).train()
    torch.manual_seed(100 + rank)  # different data shards
    x, t = torch.randn(3, *ctor[4], device=dev), torch.randn(3, *ctor[5], device=dev)

    def grads():
        m.zero_grad()
        m.train_loss(x, t).mean().backward()
        return {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    local = grads()  # no reducer attached: this rank's own gradients
    red = D.GradientAllReducer(m, dist)
    m.zero_grad()
    m.train_loss(x, t).mean().backward()
    red()
    plain = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    ov = D.OverlappedGradientReducer(m, dist)
    overlapped = grads()
    nbytes = ov.bytes_last
    ov.detach()
    # reference: mean over ranks of the local gradients, gathered explicitly (weight-gradient reductions use float
    # atomics, so two backward calls agree to ~1e-6 relative, not bit for bit)
    ok_plain, ok_ov = True, True
    for k, g in local.items():
        if k.startswith("convs.") and k.endswith(".bias"):
            continue  # zero-mean rounding noise (BatchNorm removes the mean), different on every backward call
        v = torch.view_as_real(g) if g.is_complex() else g
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v.contiguous())
        mean = sum(parts) / world
        scale = float(mean.abs().max()) + 1e-12
        a = torch.view_as_real(plain[k]) if plain[k].is_complex() else plain[k]
        b = torch.view_as_real(overlapped[k]) if overlapped[k].is_complex() else overlapped[k]
        ok_plain &= float((a - mean).abs().max()) <= 1e-4 * scale
        ok_ov &= float((b - mean).abs().max()) <= 1e-4 * scale
    differs = any(float((local[k] - plain[k]).abs().max()) > 0 for k in local if k.endswith("weights1"))
    dist.barrier()
    out.put((rank, ok_plain, ok_ov, differs, nbytes))
    dist.destroy_process_group()


def test_overlapped_allreduce_equals_plain_and_mean():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_plain, ok_ov, differs, nbytes in res:
        assert ok_plain and ok_ov, (rank, ok_plain, ok_ov)
        assert differs  # the shards really had different gradients
        assert nbytes > 0
