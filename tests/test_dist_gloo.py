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
