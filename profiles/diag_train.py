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
            rec["loss"] = float(loss)
            rec["nonfinite_grads"] = nonfinite((k, p.grad) for k, p in m.named_parameters())
        else:
            if phase == "overlap_sync":
                os.environ["B200FNO_REDUCER_SYNC"] = "1"
            overlap = phase.startswith("overlap") and dist is not None
            red = D.OverlappedGradientReducer(m, dist) if overlap else D.GradientAllReducer(m, dist)
            red.sync_parameters(0)
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            for i in range(args.steps):
                opt.zero_grad()
                loss = m.train_loss(x, t).mean()
                loss.backward()
                if not overlap:
                    red()
                bad_g = nonfinite((k, p.grad) for k, p in m.named_parameters())
                opt.step()
                bad_p = nonfinite(m.named_parameters())
                bad_b = nonfinite(m.named_buffers())
                st = {"i": i, "loss": float(loss), "bad_grads": bad_g[:6], "bad_params": bad_p[:6], "bad_buffers": bad_b[:6],
                      "same_params": same_across_ranks(m)}
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
