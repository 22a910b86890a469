#!/usr/bin/env python
"""Diagnostic for the multi-GPU training path (run under torchrun, 1 process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/diag_train.py --steps 12

Phases (each from the same initial weights, the same per-rank data as bench_train.py):
  poison    one backward with NaN-prefilled gradient buffers: any element the backward never writes stays NaN
  plain     GradientAllReducer (all-reduce after the backward), K Adam steps
  overlap   OverlappedGradientReducer (all-reduce under the backward), K Adam steps
  overlap_sync   the same with a device synchronize between the backward and the reductions (race probe)
Per step it checks: loss finite, every gradient finite, every parameter finite, parameters identical across ranks.
Prints one JSON line per phase on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_state  # noqa: E402
from bench_train import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--workload", default="fno2d_fsi_64x64_train")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--phases", default="poison,plain,overlap,overlap_sync")
    ap.add_argument("--init", default="bench", choices=["bench", "default"],
                    help="bench: bench.build_state (randomised BN); default: the module's own init under seed 0 and the "
                         "data of tests/test_gpu_multi.py")
    args = ap.parse_args()
    import realpdebench_b200 as R
    from realpdebench_b200 import dist as D
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = D.init("nccl", dev)
    ndim, modes, L, width, s_in, s_out, B = WORKLOADS[args.workload]
    B = args.batch or B
    if args.init == "default":
        torch.manual_seed(0)
        m0 = (R.FNO3d(*modes, L, width, s_in, s_out) if ndim == 3 else R.FNO2d(*modes, L, width, s_in, s_out))
        sd = {k: v.clone() for k, v in m0.state_dict().items()}
        g = torch.Generator().manual_seed(100 + rank)
        x, t = torch.randn(B, *s_in, generator=g).to(dev), torch.randn(B, *s_out, generator=g).to(dev)
    else:
        sd = build_state(ndim, modes, L, width, s_in, s_out)
        torch.manual_seed(1234 + rank)
        x, t = torch.randn(B, *s_in, device=dev), torch.randn(B, *s_out, device=dev)

    def fresh():
        m = (R.FNO3d(*modes, L, width, s_in, s_out) if ndim == 3 else R.FNO2d(*modes, L, width, s_in, s_out))
        m.load_state_dict(sd)
        return m.to(dev).train()

    def nonfinite(named):
        bad = []
        for k, v in named:
            if v is None:
                bad.append((k, "none"))
                continue
            r = torch.view_as_real(v) if v.is_complex() else v
            n = int((~torch.isfinite(r)).sum())
            if n:
                bad.append((k, n))
        return bad

    def same_across_ranks(m):
        if dist is None:
            return True
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        for p in m.parameters():
            r = (torch.view_as_real(p) if p.is_complex() else p).detach().double()
            acc[0] += r.sum()
            acc[1] += (r * r).sum()
        parts = [torch.empty_like(acc) for _ in range(world)]
        dist.all_gather(parts, acc)
        return all(bool(torch.equal(parts[0], q)) for q in parts[1:])

    for phase in args.phases.split(","):
        rec = {"phase": phase, "world": world, "rank": rank, "steps": []}
        os.environ.pop("B200FNO_POISON_GRADS", None)
        os.environ.pop("B200FNO_REDUCER_SYNC", None)
        m = fresh()
        if phase == "poison":
            os.environ["B200FNO_POISON_GRADS"] = "1"
            m.zero_grad()
            loss = m.train_loss(x, t).mean()
            loss.backward()
            torch.cuda.synchronize()
            rec["loss"] = float(loss.detach())
            rec["nonfinite_grads"] = nonfinite((k, p.grad) for k, p in m.named_parameters())
        else:
            if phase == "overlap_sync":
                os.environ["B200FNO_REDUCER_SYNC"] = "1"
            os.environ["B200FNO_REDUCER_MODE"] = phase.split("_", 1)[1] if phase.startswith("overlap_") else ""
            contend = phase == "contend"
            if contend:
                cs = torch.cuda.Stream(device=dev)
                ca = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
            os.environ["B200FNO_REDUCER_SNAPSHOT"] = "1"
            overlap = phase.startswith("overlap") and dist is not None
            red = D.OverlappedGradientReducer(m, dist) if overlap else D.GradientAllReducer(m, dist)
            red.sync_parameters(0)
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            for i in range(args.steps):
                opt.zero_grad()
                loss = m.train_loss(x, t).mean()
                if contend and i > 0:  # SM contention without NCCL: big GEMMs on a side stream under the backward
                    with torch.cuda.stream(cs):
                        for _ in range(12):
                            ca @ ca
                loss.backward()
                if not overlap:
                    red()
                bad_g = nonfinite((k, p.grad) for k, p in m.named_parameters())
                where = {}
                if os.environ.get("B200FNO_DEBUG_FINITE"):
                    from realpdebench_b200 import _capi
                    where["first_nonfinite_stage"] = int(_capi.lib().b200fno_debug_first_nonfinite(m.engine._plan))
                if getattr(red, "debug_snapshot", None):
                    where["bad_in_local_snapshot"] = [k for k, _ in nonfinite(red.debug_snapshot.items())]
                if bad_g:  # which elements, and were they already bad before the all-reduce?
                    snap = getattr(red, "debug_snapshot", {})
                    for k, _ in bad_g[:4]:
                        gk = dict(m.named_parameters())[k].grad
                        r = torch.view_as_real(gk) if gk.is_complex() else gk
                        idx = (~torch.isfinite(r)).nonzero()
                        where[k] = {"shape": list(r.shape), "first": idx[:6].tolist(), "last": idx[-3:].tolist()}
                        if k in snap:
                            sr = torch.view_as_real(snap[k]) if snap[k].is_complex() else snap[k]
                            where[k]["bad_in_local_snapshot"] = int((~torch.isfinite(sr)).sum())
                opt.step()
                bad_p = nonfinite(m.named_parameters())
                bad_b = nonfinite(m.named_buffers())
                st = {"i": i, "loss": float(loss.detach()), "bad_grads": bad_g[:6], "bad_params": bad_p[:6], "bad_buffers": bad_b[:6],
                      "same_params": same_across_ranks(m), "where": where}
                rec["steps"].append(st)
                if bad_g or bad_p or st["loss"] != st["loss"]:
                    break
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, rec)
        else:
            gathered = [rec]
        if rank == 0:
            for r in gathered:
                print(json.dumps(r), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
