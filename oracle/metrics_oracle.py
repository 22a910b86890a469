"""CPU oracle for the evaluation metrics.  TEST INFRASTRUCTURE ONLY.

Vectorised restatement of ``realpdebench/utils/metrics.py:24-131`` (``eval_metrics``) and ``:15-22``
(``kinetic_energy``): same formulas and chunking, with the two Python triple loops over (t/2, h/2, w/2) wavenumbers
(:75-81, :93-99) replaced by one index_add over a precomputed radial-bin table.  Pinned against the reference function
itself, executed in the build container (``tests/golden/make_golden.py metrics`` -> ``tests/golden/metrics.pt``).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def radial_bins(t: int, h: int, w: int) -> torch.Tensor:
    """it[i,j,k] = floor(sqrt(i^2+j^2+k^2)) for i<t//2, j<h//2, k<w//2 (metrics.py:75-78); -1 where it > nb-1 (:79-80)."""
    nb = min(t // 2, h // 2, w // 2)
    i, j, k = np.meshgrid(np.arange(t // 2), np.arange(h // 2), np.arange(w // 2), indexing="ij")
    it = np.floor(np.sqrt((i * i + j * j + k * k).astype(np.float64))).astype(np.int64)
    it[it > nb - 1] = -1
    return torch.from_numpy(it)


def _spectrum(x_F: torch.Tensor, it: torch.Tensor, nb: int) -> torch.Tensor:
    """sum of |x_F|^2 per radial bin: [b, nb, c] (metrics.py:74-81)."""
    b, c = x_F.shape[0], x_F.shape[-1]
    p = (torch.abs(x_F[:, :it.shape[0], :it.shape[1], :it.shape[2]]) ** 2).reshape(b, -1, c)
    flat = it.reshape(-1)
    keep = flat >= 0
    out = torch.zeros(b, nb, c, dtype=p.dtype)
    out.index_add_(1, flat[keep], p[:, keep])
    return out


def kinetic_energy(x: torch.Tensor) -> torch.Tensor:
    """metrics.py:15-22."""
    u = ((x[..., 0] - x[..., 0].mean(dim=1, keepdim=True)) ** 2).mean(1)
    v = ((x[..., 1] - x[..., 1].mean(dim=1, keepdim=True)) ** 2).mean(1)
    return 0.5 * (u + v)


NAMES = ("rmse", "mae", "rel_l2_error", "r2", "ke_error", "f_error", "low_f_error", "mid_f_error", "high_f_error",
         "rel_low_f_error", "rel_mid_f_error", "rel_high_f_error", "freq_error")


def eval_metrics(pred: torch.Tensor, target: torch.Tensor, c: int, batch_size=None):
    """Returns the 13 scalars of metrics.py:126-131 (as a tuple of 0-d tensors, in that order)."""
    pred_all, target_all = pred[..., :c], target[..., :c]
    b, t, h, w, c = target_all.size()
    if batch_size is None:
        batch_size = pred.shape[0]
    nb = min(t // 2, h // 2, w // 2)
    it = radial_bins(t, h, w)
    i_low, i_high = int(np.round(nb / 3)), int(np.round(nb * 2 / 3))
    rows = []
    for s in range(0, pred.shape[0], batch_size):
        p, g = pred_all[s:s + batch_size], target_all[s:s + batch_size]
        bb = p.shape[0]
        rmse = torch.sqrt(torch.mean((p - g) ** 2))
        mae = torch.mean(torch.abs(p - g))
        rel = torch.mean(torch.norm(p.reshape(bb, -1) - g.reshape(bb, -1), dim=1) / torch.norm(g.reshape(bb, -1), dim=1))
        r2 = 1 - torch.sum((p - g) ** 2) / torch.sum((g - g.mean(0, keepdim=True)) ** 2)
        ke = torch.tensor(0.) if c < 2 else (kinetic_energy(p) - kinetic_energy(g)).abs().mean()
        p_F, g_F = torch.fft.fftn(p, dim=[1, 2, 3]), torch.fft.fftn(g, dim=[1, 2, 3])
        err = torch.sqrt(torch.mean(_spectrum(p_F - g_F, it, nb), dim=0)) / (t * h * w)
        nrm = torch.sqrt(torch.mean(_spectrum(g_F, it, nb), dim=0)) / (t * h * w)
        ratio = err / nrm
        sp, sg = torch.sum(p, dim=[2, 3, 4]), torch.sum(g, dim=[2, 3, 4])
        freq = torch.mean(torch.abs(torch.fft.fftn(sp, dim=1) - torch.fft.fftn(sg, dim=1)))
        rows.append(torch.stack([rmse, mae, rel, r2, ke, err.mean(), err[:i_low].mean(), err[i_low:i_high].mean(),
                                 err[i_high:].mean(), ratio[:i_low].mean(), ratio[i_low:i_high].mean(),
                                 ratio[i_high:].mean(), freq]).float())
    m = torch.stack(rows).mean(0)
    return tuple(m[i] for i in range(13))
