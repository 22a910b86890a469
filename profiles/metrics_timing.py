"""Time realpdebench_b200.metrics.eval_metrics on one chunk at the cylinder evaluation shape
(8 samples x 200 predicted frames x 64 x 128, 2 of 3 channels) with CUDA events: device-resident and from host tensors."""
import json
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from realpdebench_b200.metrics import eval_metrics  # noqa: E402

torch.manual_seed(0)
shape = (8, 200, 64, 128, 3)
p, g = torch.randn(*shape).pin_memory(), torch.randn(*shape).pin_memory()
pd, gd = p.cuda(), g.cuda()
for _ in range(3):
    eval_metrics(pd, gd, 2, 8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = eval_metrics(pd, gd, 2, 8)
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter()
for _ in range(3):
    out_h = eval_metrics(p, g, 2, 8)
    float(out_h[0])
host_ms = (time.perf_counter() - t0) * 1e3 / 3
print(json.dumps({"shape": list(shape), "c": 2, "device_resident_ms": dev_ms, "from_pinned_host_ms": host_ms,
                  "bytes_h2d": 2 * p.numel() * 4, "rmse": float(out[0]), "f_error": float(out[5])}))
