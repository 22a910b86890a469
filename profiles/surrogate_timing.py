"""Time realpdebench_b200.materialize_surrogate (data/generate_surrogate_data.py:58-88 for one trajectory) host to host:
the script's model (modes (4,16,16), width 64, 4 layers, 10-frame windows of 128 x 128 x 17 -> 1 channel), a synthetic
trajectory of 1001 frames x 128 x 128 x 15 (984 MB), 50 windows per forward as in the script.  Wall-clock around the
call (it ends with the last device->host copy); the device-resident forward of one chunk is `python bench.py --workload
fno3d_surrogate_128x128_c17_forward`."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realpdebench_b200 as R  # noqa: E402
from bench import GaussianStats, build_state, synthetic_stats  # noqa: E402

step, batch_size, n = 10, 50, 1001
s_in, s_out = (step, 128, 128, 17), (step, 128, 128, 1)
sd = build_state(3, (4, 16, 16), 4, 64, s_in, s_out)
model = R.FNO3d(4, 16, 16, 4, 64, s_in, s_out)
model.load_state_dict(sd)
model = model.cuda().eval()
norm = GaussianStats(torch.device("cuda"), **synthetic_stats(17, 1))
rng = np.random.default_rng(0)
traj = rng.standard_normal((n, 128, 128, 15), dtype=np.float32)
times = []
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pred = R.materialize_surrogate(model, norm, traj, 40, 0.85, step=step, batch_size=batch_size)
    times.append(time.perf_counter() - t0)
ms = 1e3 * min(times[1:])
print(json.dumps({"frames": n, "pred_shape": list(pred.shape), "windows_per_forward": batch_size,
                  "host_to_host_ms": ms, "field_points_per_s": pred.size / (ms * 1e-3),
                  "h2d_bytes": traj.nbytes + 10 * 128 * 128 * 15 * 4, "d2h_bytes": (pred.shape[0] + step - 1) * 128 * 128 * 4,
                  "engine_impl": model.engine.resolved_impl(), "pred_mean": float(pred.mean())}))
