"""Measure tcgen05.mma kind::tf32 cycles per instruction (M=128, K=8) on the B200: python profiles/mma_rate.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realpdebench_b200 import _capi  # noqa: E402

out = {}
buf = torch.zeros(1, dtype=torch.int64, device="cuda")
for N in (16, 64, 128, 256):
    for ts in (0, 1):
        for nacc in (1, 2, 4):
            for acc in (0, 1):
                if nacc * N > 448:
                    continue
                res = []
                for iters in (256, 2048):
                    _capi.check(_capi.lib().b200fno_selftest_mma_rate(N, iters, ts, nacc, acc, buf.data_ptr(), 0))
                    torch.cuda.synchronize()
                    res.append(int(buf.item()))
                out[f"N{N}_{'ts' if ts else 'ss'}_nacc{nacc}_{'acc' if acc else 'ovw'}"] = (res[1] - res[0]) / (2048 - 256)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mma_rate.json", "w"), indent=1)
