"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: batch sharding and the timing reduction
bench.py uses (whole-job units / max-over-ranks time).  The sharded oracle forward must equal the
unsharded one because the path has no inter-sample coupling at inference."""
import os
import socket

import torch
import torch.multiprocessing as mp

from oracle import fno_oracle as O
from realpdebench_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist = D.init("gloo")
    torch.manual_seed(0)
    s = (4, 8, 12, 3)
    sd = O.init_state(3, (2, 3, 4), 2, 8, s, s)
    O.randomize_bn(sd)
    torch.manual_seed(1)
    x = torch.randn(5, *s)  # global batch 5 over 2 ranks: 3 + 2
    lo, hi = D.shard_range(5, rank, world)
    y = O.fno3d_forward(sd, x[lo:hi], s)
    full = O.fno3d_forward(sd, x, s)
    ok = torch.allclose(y, full[lo:hi], atol=1e-6)
    rate, ms = D.aggregate_throughput(units_this_rank=float(hi - lo), ms_this_rank=10.0 * (rank + 1), dist=dist)
    dist.barrier()
    out.put((rank, lo, hi, ok, rate, ms))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]  # contiguous, disjoint, covering
    assert all(r[3] for r in res)
    for r in res:  # both ranks agree: 5 units / max(10, 20) ms
        assert abs(r[5] - 20.0) < 1e-9 and abs(r[4] - 5 / 0.020) < 1e-6


def test_shard_range_properties():
    for gb in (1, 7, 8, 64):
        for w in (1, 2, 4, 8):
            spans = [D.shard_range(gb, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------- data-parallel training: gradient all-reduce
class _Params(torch.nn.Module):
    """Parameter container with the reference FNO3d names/dtypes (complex64 spectral weights included)."""

    def __init__(self, sd):
        super().__init__()
        self.names = [k for k in sd if O.is_param(k)]
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(sd[k].clone()) for k in self.names])
        self.register_buffer("running", torch.zeros(3))


def _grads_of_shard(sd, x, t, s):
    _, g, _ = O.train_loss_and_grads(3, {k: v.clone() for k, v in sd.items()}, x, t, s)
    return g


def _train_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist = D.init("gloo")
    torch.manual_seed(0)
    s = (4, 8, 12, 3)
    sd = O.init_state(3, (2, 3, 4), 2, 8, s, s)
    torch.manual_seed(1)
    x, t = torch.randn(4, *s), torch.randn(4, *s)
    m = _Params(sd)
    if rank == 1:  # diverged replica: sync_parameters must restore rank 0's values
        with torch.no_grad():
            for p in m.ps:
                p.mul_(0.5)
            m.running.fill_(7.0)
    red = D.GradientAllReducer(m, dist, bucket_bytes=4096)  # small buckets: several messages, some params split off
    red.sync_parameters(0)
    same = all(torch.equal(p.data, sd[k]) for k, p in zip(m.names, m.ps)) and float(m.running.sum()) == 0.0
    # the broadcast writes through .data (no version bump): the engine's packed-weight key must be invalidated explicitly
    same = same and getattr(m, "_stats_epoch", 0) == 1
    red.sync_buffers(0)
    same = same and m._stats_epoch == 2
    lo, hi = D.shard_range(4, rank, world)
    mine = _grads_of_shard(sd, x[lo:hi], t[lo:hi], s)
    for k, p in zip(m.names, m.ps):
        p.grad = mine[k].clone()
    nbytes = red()
    both = [_grads_of_shard(sd, x[a:b], t[a:b], s) for a, b in (D.shard_range(4, r, world) for r in range(world))]
    ok = all(torch.allclose(p.grad, (both[0][k] + both[1][k]) / 2, rtol=1e-6, atol=1e-9) for k, p in zip(m.names, m.ps))
    dist.barrier()
    out.put((rank, same, ok, nbytes, len(red.buckets)))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_real = None
    for rank, same, ok, nbytes, nbuckets in res:
        assert same and ok and nbuckets > 1
        n_real = nbytes if n_real is None else n_real
        assert nbytes == n_real and nbytes > 0
