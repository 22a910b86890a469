#!/usr/bin/env python
"""Eager (PDL-chained launches) vs CUDA-graph replay of model.rollout at small batch (VERDICT r01 "smaller" item)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, build_model  # noqa: E402


def main():
    import realpdebench_b200 as R
    dev = torch.device("cuda:0")
    out = []
    for wl, B in (("fno3d_combustion_128x128x64_rollout10", 1), ("fno3d_cylinder_64x128_rollout10", 1),
                  ("fno2d_cylinder_256x512_rollout20", 1), ("fno2d_cylinder_256x512_rollout20", 8)):
        ndim, modes, L, width, s_in, s_out, _, n_auto = WORKLOADS[wl][:8]
        m = build_model(R, ndim, modes, L, width, s_in, s_out).to(dev).eval()
        x0 = torch.randn(B, *s_in, device=dev)
        a, b = torch.ones(s_out[-1], device=dev), torch.zeros(s_out[-1], device=dev)
        res = {"workload": wl, "batch": B, "n_steps": n_auto}
        for graph in (False, True):
            for _ in range(3):
                y = m.rollout(x0, a, b, n_auto, graph=graph)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                y = m.rollout(x0, a, b, n_auto, graph=graph)
            e1.record()
            torch.cuda.synchronize()
            res["graph_ms" if graph else "eager_ms"] = e0.elapsed_time(e1) / 10
        res["equal"] = bool(torch.equal(m.rollout(x0, a, b, n_auto), m.rollout(x0, a, b, n_auto, graph=True)))
        out.append(res)
        del m
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
