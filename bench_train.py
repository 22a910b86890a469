#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[2] / SURVEY.md 8d row C3): FNO-2D on the fsi geometry
(64x64 grid, 20 frames x 3 channels, modes (16,16), width 128, 4 layers), batch 32 per GPU, one step of
train.py:321-334 = zero_grad, train-mode forward, MSE, backward, [gradient all-reduce], Adam step.

    python bench_train.py --gpus N --steps K --warmup W
    python bench_train.py --impl reference        # the oracle's train step on the host cores (bounded sample)

This is NOT the headline bench (bench.py is); it prints one JSON line in the same shape.  fp32: the
autocast-bf16 variant of C3 is not built.  One field-point = one predicted (b,t,h,w).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import ClockSampler, build_model, build_state  # noqa: E402

WORKLOADS = {
    # name: (ndim, modes, n_layers, width, shape_in, shape_out, batch per GPU)
    "fno2d_fsi_64x64_train": (2, (16, 16), 4, 128, (20, 64, 64, 3), (20, 64, 64, 3), 32),
    "fno3d_cylinder_64x128_train": (3, (4, 12, 16), 4, 64, (20, 64, 128, 3), (20, 64, 128, 3), 4),
}
METRIC = "fno_train_field_points_per_sec"


def config(wl, B, world):
    ndim, modes, L, width, s_in, s_out, _ = WORKLOADS[wl]
    return {"workload": wl, "operator": f"fno{ndim}d", "modes": list(modes), "width": width, "n_layers": L,
            "shape_in": list(s_in), "shape_out": list(s_out), "batch_per_gpu": B, "global_batch": B * world,
            "optimizer": "torch.optim.Adam(lr=1e-3) as train.py:290", "parallelism":
            f"batch-sharded x{world}, one gradient all-reduce per step (overlapped with the backward pass up to 4 ranks: "
            "realpdebench_b200.dist.OverlappedGradientReducer; after it at 8: GradientAllReducer)"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import fno_oracle as O
    ndim, modes, L, width, s_in, s_out, B = WORKLOADS[args.workload]
    B = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = build_state(ndim, modes, L, width, s_in, s_out)
    torch.manual_seed(1234)
    batches = [(torch.randn(B, *s_in), torch.randn(B, *s_out))]
    times = []
    for i in range(1 + max(1, args.steps)):
        t0 = time.perf_counter()
        O.train_steps(ndim, sd, batches, s_out, 1e-3)
        if i >= 1:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    pts = B * s_out[0] * s_out[1] * s_out[2]
    v = pts / (ms * 1e-3)
    from bench import cpu_model_name
    base = {"value": v, "unit": "field-points/s", "cores": cores, "cpu_model": cpu_model_name(), "kind": "port",
            "sample": f"oracle/fno_oracle.py train_steps (fwd + autograd bwd + Adam), batch {B} of {args.workload}, "
                      f"mean of {len(times)} step(s) after 1 warm-up, torch {torch.__version__} CPU {cores} threads"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "field-points/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args.workload, B, 1),
                      "cpu_baseline": base, "gpu_launches": 0}), flush=True)


def run_engine(args):
    import realpdebench_b200 as R
    from realpdebench_b200 import _capi
    from realpdebench_b200 import dist as D

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench_train.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = D.init("nccl", dev)
    ndim, modes, L, width, s_in, s_out, B = WORKLOADS[args.workload]
    B = args.batch or B
    model = build_model(R, ndim, modes, L, width, s_in, s_out).to(dev).train()
    if args.dtype == "bf16":  # BASELINE config C3 as named: the autocast arithmetic (GEMM operands in bf16, fp32 accumulate)
        model.set_compute("bf16")
    if args.fused_adam:  # same updates in one pass per parameter (realpdebench_b200/optim.py)
        from realpdebench_b200.optim import FusedAdam
        optimizer = FusedAdam(model.parameters(), lr=1e-3)
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)  # train.py:290
    # N > 1: the all-reduce runs under the backward pass (events recorded by b200fno_train_backward); --no-overlap
    # falls back to one bucketed all-reduce after loss.backward()
    # default: overlapped up to 4 ranks, after-backward at 8 - measured on 8 B200 (profiles/r02_train_bench_n8*.json):
    # 11.05 ms plain vs 12.89 ms overlapped on a capped 8-CTA communicator vs 19.2 ms overlapped on the stock one
    # (the collectives of 8 ranks wait on each other while holding SMs of the backward pass); --overlap forces it
    overlap = dist is not None and not args.no_overlap and (world <= 4 or args.overlap)
    reducer = D.OverlappedGradientReducer(model, dist) if overlap else D.GradientAllReducer(model, dist)
    reducer.sync_parameters(0)
    torch.manual_seed(1234 + rank)
    x, t = torch.randn(B, *s_in, device=dev), torch.randn(B, *s_out, device=dev)
    xh, th = x.cpu().pin_memory(), t.cpu().pin_memory()
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for k in ("fwd", "bwd", "allreduce", "adam")}

    def step(xb, tb, timed=False):
        optimizer.zero_grad()
        if timed:
            ev["fwd"][0].record()
        loss = model.train_loss(xb, tb).mean()  # train.py:328
        if timed:
            ev["fwd"][1].record(), ev["bwd"][0].record()
        loss.backward()
        if timed:
            ev["bwd"][1].record(), ev["allreduce"][0].record()
        nbytes = reducer.bytes_last if overlap else reducer()
        if timed:
            ev["allreduce"][1].record(), ev["adam"][0].record()
        optimizer.step()
        if timed:
            ev["adam"][1].record()
        return loss, nbytes

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(x, t)
    barrier()
    _capi.lib().b200fno_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            loss, nbytes = step(x, t)
        e1.record()
        barrier()
    launches = int(_capi.lib().b200fno_launch_count())
    ms_step = D.max_over_ranks(e0.elapsed_time(e1), dist, dev) / args.steps
    pts = B * s_out[0] * s_out[1] * s_out[2]
    value = world * pts / (ms_step * 1e-3)
    step(x, t, timed=True)
    torch.cuda.synchronize()
    phases = {k: round(a.elapsed_time(b), 4) for k, (a, b) in ev.items()}
    # end to end: host batch -> device every step, loss read back (train.py:324-335)
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        l, _ = step(xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True))
        l.item()
    barrier()
    e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e, dist, dev)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "field-points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config(args.workload, B, world),
            "impl": "b200fno", "samples_per_sec": world * B / (ms_step * 1e-3),
            "e2e": {"value": world * pts / (e2e_ms * 1e-3), "unit": "field-points/s",
                    "h2d_bytes_per_step": (xh.numel() + th.numel()) * 4, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "clocks": clocks.summary(), "phases_ms": phases,
            "allreduce_bytes_per_step": nbytes, "allreduce_overlapped": overlap,
            "reducer_max_ctas": getattr(reducer, "max_ctas", None), "fused_adam": bool(args.fused_adam), "loss": float(loss.detach())}), flush=True)
    # a step that produced a non-finite loss or parameter is not a measurement: fail loudly on every rank
    finite = bool(torch.isfinite(loss.detach())) and all(
        bool(torch.isfinite(torch.view_as_real(p) if p.is_complex() else p).all()) for p in model.parameters())
    if dist is not None:
        flag = torch.tensor([0.0 if finite else 1.0], device=dev)
        dist.all_reduce(flag)
        finite = float(flag.item()) == 0.0
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench_train.py: non-finite loss or parameter after the timed steps - the numbers above are void")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200fno", choices=["b200fno", "reference"])
    ap.add_argument("--workload", default="fno2d_fsi_64x64_train", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16 = torch.autocast(bfloat16) arithmetic of the reference for every Linear / Conv GEMM of the "
                         "step (forward + backward); FFMA kernels either way, so this is the numerics, not a speed-up")
    ap.add_argument("--fused-adam", action="store_true", help="realpdebench_b200.optim.FusedAdam instead of torch's Adam")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after the backward pass instead of under it")
    ap.add_argument("--overlap", action="store_true", help="force the overlapped all-reduce also beyond 4 ranks")
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the bounded CPU sample (--impl reference)")
    args = ap.parse_args()
    (run_reference if args.impl == "reference" else run_engine)(args)


if __name__ == "__main__":
    main()
