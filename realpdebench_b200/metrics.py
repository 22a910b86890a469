"""``eval_metrics`` with the reference's signature (realpdebench/utils/metrics.py:24-131) on the GPU.

The reference computes the 13 evaluation scalars on whatever device the tensors live on, binning the Fourier-space
error with two Python triple loops over (t/2, h/2, w/2) wavenumbers (:75-81, :93-99) - the dominant cost of ``eval.py``
once the rollout is fast.  Here each chunk is one C-ABI call (``b200fno_eval_metrics``): truncated
DFTs (only wavenumbers below ``min(t,h,w)/2`` are ever binned), radial binning and all reductions run in CUDA kernels.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import check

NAMES = ("rmse", "mae", "rel_l2_error", "r2", "ke_error", "f_error", "low_f_error", "mid_f_error", "high_f_error",
         "rel_low_f_error", "rel_mid_f_error", "rel_high_f_error", "freq_error")


def eval_metrics(pred: torch.Tensor, target: torch.Tensor, c: int, batch_size=None, device=None):
    """pred, target: [b, t, h, w, c'] float tensors, ``c`` channels evaluated.  Returns the 13 scalars of
    metrics.py:126-131 (0-d tensors on the device), each the mean over the ``batch_size`` chunks like the reference.

    The tensors may live on the host (``eval.py:342-343`` collects ``pred.cpu()``): each chunk is copied to ``device``
    (default: the tensors' CUDA device, else the current one) right before its C-ABI call, so the concatenated
    prediction never has to fit in HBM at once.  ``batch_size=None`` is ONE chunk like the reference (r2 and the
    spectra depend on the chunking, so it cannot be split behind the caller's back)."""
    if pred.shape != target.shape or pred.dim() != 5:
        raise RuntimeError(f"b200fno: eval_metrics expects two [b,t,h,w,c] tensors, got {tuple(pred.shape)}, {tuple(target.shape)}")
    if device is None:
        device = pred.device if pred.is_cuda else (target.device if target.is_cuda else None)
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("b200fno: eval_metrics runs on CUDA only (no CPU fallback) and no CUDA device is available")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    b, t, h, w, ct = pred.shape
    if batch_size is None:
        batch_size = b
    L = _capi.lib()
    need = L.b200fno_metrics_workspace_bytes(min(batch_size, b), t, h, w, ct, c)
    if need == 0:
        check(-1)
    ws = torch.empty(need, dtype=torch.uint8, device=device)
    rows = []
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        for s in range(0, b, batch_size):
            p = pred[s:s + batch_size].to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
            g = target[s:s + batch_size].to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
            out = torch.empty(13, dtype=torch.float32, device=device)
            check(L.b200fno_eval_metrics(p.data_ptr(), g.data_ptr(), p.shape[0], t, h, w, ct, c, ws.data_ptr(), need,
                                         out.data_ptr(), stream))
            rows.append(out)
    m = torch.stack(rows).mean(0)
    return tuple(m[i] for i in range(13))
