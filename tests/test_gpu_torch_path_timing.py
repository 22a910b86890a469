"""The reference's own arithmetic (torch ops: cuFFT rfft2/irfft2, einsum, conv, batch_norm, gelu) run ON THE
SAME B200 through the oracle restatement, timed beside the engine on BASELINE config C2 (FNO-2D 256x512,
batch 8).  This is the "what if one just ran the PyTorch model on the GPU" number that DESIGN.md section 2
argues against; the result is written to gpurun_out/torch_gpu_path.json (copied to profiles/ per round)."""
import json
import os

import pytest
import torch

from oracle import fno_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_engine_beats_torch_cufft_path_on_c2():
    import realpdebench_b200 as R
    dev = torch.device("cuda:0")
    s = (20, 256, 512, 3)
    modes, L, width, B, n_auto = (12, 16), 4, 64, 8, 4
    torch.manual_seed(0)
    sd = O.init_state(2, modes, L, width, s, s)
    O.randomize_bn(sd)
    m = R.FNO2d(*modes, L, width, s, s)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    norm = O.synthetic_normalizer(3, 3)
    for k, v in list(vars(norm).items()):
        if torch.is_tensor(v):
            setattr(norm, k, v.to(dev))
    norm.device = dev
    torch.manual_seed(1)
    x = torch.randn(B, *s, device=dev)
    tgt = torch.randn(B, n_auto * s[0], *s[1:], device=dev)
    a, b = R.rollout_affine(norm, 3, 3, dev)
    x0 = norm.preprocess(x, tgt[:, :1])[0].contiguous()
    res = {}
    with torch.no_grad():
        for tf32 in (False, True):  # torch defaults: cudnn TF32 on, matmul TF32 off; also time everything-fp32
            torch.backends.cudnn.allow_tf32 = tf32
            ms = _time(lambda: O.rollout(lambda t: O.fno2d_forward(sd_dev, t, s), norm, x, tgt, n_auto), 3)
            res["torch_gpu_ms_per_step_tf32conv" if tf32 else "torch_gpu_ms_per_step_fp32"] = ms / n_auto
        ms_eng = _time(lambda: m.rollout(x0, a, b, n_auto), 5) / n_auto
        # same results (the torch path with fp32 convolutions is the oracle itself, on another device)
        torch.backends.cudnn.allow_tf32 = False
        pred_t = O.rollout(lambda t: O.fno2d_forward(sd_dev, t, s), norm, x, tgt[:, :s[0]], 1)[3][1]
        pred_e = m.rollout(x0, a, b, 1)
        err = O.rel_l2(pred_e, pred_t)
    res.update(engine_ms_per_step=ms_eng, batch=B, workload="fno2d_cylinder_256x512 (C2), per autoregressive step",
               rel_l2_engine_vs_torch_gpu=err, speedup_vs_torch_gpu_fp32=res["torch_gpu_ms_per_step_fp32"] / ms_eng,
               torch=torch.__version__)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "torch_gpu_path.json"), "w") as f:
        json.dump(res, f)
    print(json.dumps(res))
    assert err < 1e-5
    assert ms_eng < res["torch_gpu_ms_per_step_fp32"]
