"""Oracle value of bench.py's ``normalized_loss_check`` (eval.py:323) for the headline workload, rank 0.

bench.py's engine arm prints the normalised rollout loss of its seeded synthetic batch; this script computes the
same number with the CPU oracle (``oracle.fno_oracle.rollout`` = eval.py:296-326 on the oracle's FNO-2D forward)
on the very same tensors (``torch.manual_seed(1234 + rank)``, rank 0) and writes it to
``tests/golden/bench_loss_check.json``.  bench.py compares its measured value with it on every run.

    python tests/golden/make_bench_loss.py            # ~2 min on 8 cores, 12 GB of host memory
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oracle import fno_oracle as O  # noqa: E402


def main():
    out = {}
    for wl in (bench.DEFAULT_WORKLOAD,):
        ndim, modes, L, width, s_in, s_out, B, n_auto = bench.WORKLOADS[wl]
        sd = bench.build_state(ndim, modes, L, width, s_in, s_out)
        norm = O.Normalizer("gaussian", **bench.synthetic_stats(s_in[-1], s_out[-1]))
        torch.manual_seed(1234)  # rank 0 of bench.py
        x = torch.randn(B, *s_in)
        tgt = torch.randn(B, n_auto * s_out[0], *s_out[1:])
        fwd = (lambda t: O.fno3d_forward(sd, t, s_out)) if ndim == 3 else (lambda t: O.fno2d_forward(sd, t, s_out))
        losses = []
        with torch.no_grad():  # per sample: the loss is a mean over equally sized samples
            for i in range(B):
                _, _, l, _ = O.rollout(fwd, norm, x[i:i + 1], tgt[i:i + 1], n_auto)
                losses.append(l)
                print(wl, i, l, flush=True)
        out[wl] = {"normalized_loss": sum(losses) / len(losses), "per_sample": losses, "seed": 1234, "batch": B,
                   "n_autoregressive": n_auto, "torch": torch.__version__}
    with open(os.path.join(ROOT, "tests", "golden", "bench_loss_check.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
