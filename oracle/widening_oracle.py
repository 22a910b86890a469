"""CPU oracle for the callers next to the hot path.  TEST INFRASTRUCTURE ONLY.

Restated here, over plain tensors / numpy arrays:

* ``realpdebench/data/generate_surrogate_data.py:58-88``  the surrogate materialisation loop for ONE trajectory
  file (script body, not importable: hard-coded paths, 128 x 128 x 15 shapes; the shapes are taken from the array here)
* ``realpdebench/model/MWT_libs/models.py:535-585``       ``sparseKernelFT3d.forward``
* ``realpdebench/model/MWT_libs/models.py:252-295``       ``sparseKernelFT2d.forward``
* ``realpdebench/model/galerkin_transformer_libs/layers.py:1205-1257``  Galerkin ``SpectralConv3d.forward``

Parity pin: ``tests/golden/make_golden.py widening`` executes the reference source (the script lines exec'd verbatim,
the modules imported) in the build container and commits inputs / outputs as ``tests/golden/surrogate.pt`` and
``tests/golden/siblings.pt``; ``tests/test_oracle.py`` checks this file against them.

Only ``tests/`` may import this module.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .fno_oracle import Normalizer, spectral_conv2d, spectral_conv3d

Tensor = torch.Tensor


def _with_parameter_channels(window: Tensor, gas_ratio, equivalence_ratio) -> Tensor:
    """generate_surrogate_data.py:66-68 / 78-80: two constant channels appended after the measured ones."""
    gas = torch.ones_like(window[..., [0]]) * gas_ratio
    eq = torch.ones_like(window[..., [0]]) * equivalence_ratio
    return torch.cat([window, gas, eq], dim=-1)


def materialize_surrogate(model_fn: Callable[[Tensor], Tensor], norm: Normalizer, traj_numerical: np.ndarray,
                          gas_ratio, equivalence_ratio, step: int = 10, batch_size: int = 50,
                          sub_s: int = 1) -> np.ndarray:
    """generate_surrogate_data.py:58-88 for one trajectory ``traj_numerical [n, H, W, C]``.

    Windows of ``step`` frames, ``batch_size`` windows per forward, every window predicted independently (no
    autoregression); the last frame comes from one extra window over the final ``step`` frames (:76-86).
    Returns ``pred_traj`` of :88, shape ``[n_pred, H', W']``.
    """
    h, w, c = traj_numerical[:, ::sub_s, ::sub_s].shape[1:]
    pred_list = []
    with torch.no_grad():
        for i in range(0, traj_numerical.shape[0] - 1, batch_size * step):  # :63
            x = torch.tensor(traj_numerical[i:i + batch_size * step, ::sub_s, ::sub_s], dtype=torch.float) \
                .reshape(-1, step, h, w, c)  # :65
            x = _with_parameter_channels(x, gas_ratio, equivalence_ratio)
            x, _ = norm.preprocess(x, x)  # :70
            p = model_fn(x)  # :71
            _, p = norm.postprocess(p, p)  # :72
            pred_list.append(p.reshape(-1, h, w).cpu().numpy())  # :74
        x = torch.tensor(traj_numerical[-step:, ::sub_s, ::sub_s], dtype=torch.float).reshape(1, step, h, w, c)  # :77
        x = _with_parameter_channels(x, gas_ratio, equivalence_ratio)
        x, _ = norm.preprocess(x, x)
        p = model_fn(x)
        _, p = norm.postprocess(p, p)
        pred_list.append(p.reshape(-1, h, w)[[-1]].cpu().numpy())  # :86
    return np.concatenate(pred_list, axis=0)  # :88


def mwt_sparse_kernel_ft3d(x: Tensor, weights: Sequence[Tensor], modes: int, lo_w: Tensor, lo_b: Tensor) -> Tensor:
    """MWT_libs/models.py:557-585.  x: [B, Nx, Ny, T, c, k^2]; weights: 4 x complex [C, C, modes, modes, modes]."""
    B, Nx, Ny, T, c, ich = x.shape
    z = x.reshape(B, Nx, Ny, T, -1).permute(0, 4, 1, 2, 3)
    l1, l2 = min(modes, Nx // 2 + 1), min(modes, Ny // 2 + 1)  # :565-566
    z = spectral_conv3d(z, *[wt[:, :, :l1, :l2, :] for wt in weights])  # same 4 corners in the same order (:569-576)
    z = F.relu(z.permute(0, 2, 3, 4, 1))
    return F.linear(z, lo_w, lo_b).reshape(B, Nx, Ny, T, c, ich)


def mwt_sparse_kernel_ft2d(x: Tensor, weights: Sequence[Tensor], modes: int, lo_w: Tensor, lo_b: Tensor) -> Tensor:
    """MWT_libs/models.py:270-295.  x: [B, Nx, Ny, c, k^2]; weights: 2 x complex [C, C, modes, modes]."""
    B, Nx, Ny, c, ich = x.shape
    z = x.reshape(B, Nx, Ny, -1).permute(0, 3, 1, 2)
    l1, l2 = min(modes, Nx // 2 + 1), min(modes, Ny // 2 + 1)  # :277-279
    z = spectral_conv2d(z, *[wt[:, :, :l1, :l2] for wt in weights])
    z = F.relu(z.permute(0, 2, 3, 1))
    return F.linear(z, lo_w, lo_b).reshape(B, Nx, Ny, c, ich)


def galerkin_spectral_conv3d(x: Tensor, weights: Sequence[Tensor]) -> Tensor:
    """galerkin_transformer_libs/layers.py:1238-1257: the FNO operator itself (``modes1`` is named ``modes_t``)."""
    return spectral_conv3d(x, *weights)
