"""Fused Adam for the engine's training path.

``train.py:290`` builds ``torch.optim.Adam(model.parameters(), lr=args.lr)``; this class takes the same arguments
and produces the same updates (same formulas in the same order as torch's foreach implementation, defaults only:
no weight decay, no amsgrad) but touches every parameter ONCE: one CUDA kernel per tensor reads p, g, m, v and writes
p, m, v, instead of torch's ~10 multi-tensor passes.  Complex spectral weights are updated through their real views
(re and im independent, exactly what torch does).  LR schedulers work unchanged (they edit ``param_groups``).
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _capi.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("b200fno FusedAdam runs on CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                real = (lambda t: torch.view_as_real(t) if t.is_complex() else t)
                pv, gv, mv, vv = real(p.data), real(p.grad), real(st["exp_avg"]), real(st["exp_avg_sq"])
                if pv.dtype != torch.float32 or not (pv.is_contiguous() and gv.is_contiguous()):
                    raise RuntimeError("b200fno FusedAdam needs contiguous float32 / complex64 parameters and gradients")
                stream = torch.cuda.current_stream(p.device).cuda_stream
                with torch.cuda.device(p.device):
                    check(L.b200fno_adam_step(pv.data_ptr(), gv.data_ptr(), mv.data_ptr(), vv.data_ptr(), pv.numel(),
                                              float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                              int(st["step"]), stream))
                # the kernel wrote through raw pointers: bump torch's version counter (an in-place op on an empty
                # view, no kernel launch) so that the engine re-packs its weight copy on the next forward
                p.view(-1)[:0].zero_()
        return loss
