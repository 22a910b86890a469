"""CPU check of the training path's algorithm (csrc/api.cu: b200fno_train_backward): the backward of the spectral
operator is emulated in numpy as the engine runs it - the forward stage kernels applied with TRANSPOSED tables, the
per-mode mixing with conj(W)^T, the weight gradient conj(S) (x) dO scattered back to the corner tensors (overwritten
corner elements get zero) - using the library's own tables (b200fno_host_table, no device needed), and compared with
torch autograd of the oracle's SpectralConv."""
import numpy as np
import pytest
import torch

from oracle import fno_oracle as O
from realpdebench_b200 import _capi


def _tables(ndim, T, H, W, m1, m2, m3):
    tabs = [_capi.host_table(ndim, T, H, W, m1, m2, m3, k) for k in range(6)]
    (LF, ft, fh) = tabs[0]
    f64 = lambda a: a.astype(np.float64)
    return dict(LF=f64(LF), LH=f64(tabs[1][0]), LT=f64(tabs[2][0]), LTi=f64(tabs[3][0]), LHi=f64(tabs[4][0]),
                Gt=f64(tabs[5][0]), ft=ft, fh=fh)


def _corner(ndim, f_t, f_h, T, H, m1, m2):
    """(corner index, x, y) feeding kept frequency (f_t, f_h): the rule of pack.cu (later assignment wins)."""
    h_hi = f_h >= H - m2
    y = f_h - (H - m2) if h_hi else f_h
    if ndim == 3:
        t_hi = f_t >= T - m1
        return (2 if h_hi else 0) + (1 if t_hi else 0), (f_t - (T - m1) if t_hi else f_t), y
    return (1 if h_hi else 0), 0, y


def emulate_forward_backward(ndim, x, weights, dy, m1, m2, m3):
    """x, dy: [B,C,(T,)H,W] float64; weights: complex corner arrays.  Returns (y, dx, dweights)."""
    if ndim == 2:
        x, dy = x[:, :, None], dy[:, :, None]
    B, Ci, T, H, W = x.shape
    Co = weights[0].shape[1]
    tb = _tables(ndim, T, H, W, m1, m2, m3)
    ft, fh = tb["ft"], tb["fh"]
    KT, KH, K2 = len(ft), len(fh), 2 * m3
    LF, LH, LHi, Gt = tb["LF"][:K2, :W], tb["LH"][:2 * KH, :2 * H], tb["LHi"][:2 * H, :2 * KH], tb["Gt"][:W, :K2]
    LT, LTi = tb["LT"][:2 * KT, :2 * T], tb["LTi"][:2 * T, :2 * KT]
    cl = lambda a: np.transpose(a, (0, 2, 3, 4, 1))  # channels-last
    # ---- forward (api.cu: run_spectral + the inverse-W term of the layer kernel)
    A = np.einsum("mw,bthwc->bthmc", LF, cl(x)).reshape(B, T, H * 2, m3, Ci)
    Bh = np.einsum("mk,btknc->btmnc", LH, A)
    S = (np.einsum("mk,bkhnc->bmhnc", LT, Bh.reshape(B, T * 2, KH, m3, Ci)) if ndim == 3
         else Bh.reshape(B, 2 * KT, KH, m3, Ci)).reshape(B, 2, KT, KH, m3, Ci)
    Sc = S[:, 0] + 1j * S[:, 1]
    Oc = np.zeros((B, KT, KH, m3, Co), dtype=np.complex128)
    for a, f_t in enumerate(ft):
        for b_, f_h in enumerate(fh):
            c, xx, yy = _corner(ndim, f_t, f_h, T, H, m1, m2)
            wc = weights[c][:, :, xx, yy, :] if ndim == 3 else weights[c][:, :, yy, :]
            Oc[:, a, b_] = np.einsum("bzi,ioz->bzo", Sc[:, a, b_], wc)
    Or = np.stack([Oc.real, Oc.imag], axis=1)
    Ct = (np.einsum("mk,bkhnc->bmhnc", LTi, Or.reshape(B, 2 * KT, KH, m3, Co)) if ndim == 3
          else Or.reshape(B, 2 * KT, KH, m3, Co)).reshape(B, T, 2 * KH, m3, Co)
    D = np.einsum("mk,btknc->btmnc", LHi, Ct).reshape(B, T, H, K2, Co)
    y = np.einsum("wk,bthkc->bthwc", Gt, D)
    # ---- backward: the same stage kernels with transposed tables, last stage first
    dD = np.einsum("wk,bthwc->bthkc", Gt, cl(dy)).reshape(B, T, H * 2, m3, Co)      # lmul(Gt^T)
    dCt = np.einsum("mk,btmnc->btknc", LHi, dD)                                      # lmul(LHi^T)
    dOr = (np.einsum("mk,bmhnc->bkhnc", LTi, dCt.reshape(B, T * 2, KH, m3, Co)) if ndim == 3
           else dCt.reshape(B, 2 * KT, KH, m3, Co)).reshape(B, 2, KT, KH, m3, Co)    # lmul(LTi^T)
    dOc = dOr[:, 0] + 1j * dOr[:, 1]  # (d/dRe, d/dIm) pairs = torch's gradient convention for complex tensors
    dSc = np.zeros_like(Sc)
    dW = [np.zeros_like(w) for w in weights]
    for a, f_t in enumerate(ft):
        for b_, f_h in enumerate(fh):
            c, xx, yy = _corner(ndim, f_t, f_h, T, H, m1, m2)
            wc = weights[c][:, :, xx, yy, :] if ndim == 3 else weights[c][:, :, yy, :]
            dSc[:, a, b_] = np.einsum("bzo,ioz->bzi", dOc[:, a, b_], np.conj(wc))    # modes_kernel(dO, conj(W)^T)
            g = np.einsum("bzi,bzo->ioz", np.conj(Sc[:, a, b_]), dOc[:, a, b_])      # modes_wgrad_kernel
            if ndim == 3:
                dW[c][:, :, xx, yy, :] = g                                           # unpack_spectral_grad_kernel
            else:
                dW[c][:, :, yy, :] = g
    dS = np.stack([dSc.real, dSc.imag], axis=1).reshape(B, 2 * KT, KH, m3, Ci)
    dBh = (np.einsum("mk,bmhnc->bkhnc", LT, dS) if ndim == 3 else dS).reshape(B, T, 2 * KH, m3, Ci)  # lmul(LT^T)
    dA = np.einsum("mk,btmnc->btknc", LH, dBh).reshape(B, T, H, K2, Ci)              # lmul(LH^T)
    dx = np.einsum("mw,bthmc->bthwc", LF, dA)                                        # layer kernel with LF^T
    back = lambda a: np.transpose(a, (0, 4, 1, 2, 3))
    y, dx = back(y), back(dx)
    if ndim == 2:
        y, dx = y[:, :, 0], dx[:, :, 0]
    return y, dx, dW


@pytest.mark.parametrize("ndim,shape,modes,ci,co", [
    (3, (9, 10, 12), (2, 3, 4), 3, 4),
    (3, (5, 6, 7), (3, 4, 4), 2, 3),     # overlapping corners on T and H: overwritten elements must get zero gradient
    (3, (8, 7, 16), (3, 2, 9), 2, 2),    # Nyquist bin kept
    (2, (14, 18), (5, 5), 3, 4),
    (2, (9, 8), (5, 5), 2, 2),           # overlapping corners in 2-D
])
def test_backward_stage_sequence_matches_autograd(ndim, shape, modes, ci, co):
    torch.manual_seed(3)
    m = modes if ndim == 3 else (1, *modes)
    x = torch.randn(2, ci, *shape, dtype=torch.float64, requires_grad=True)
    ws = [torch.randn(ci, co, *modes, dtype=torch.cdouble, requires_grad=True) for _ in range(4 if ndim == 3 else 2)]
    dy = torch.randn(2, co, *shape, dtype=torch.float64)
    y = O.spectral_conv3d(x, *ws) if ndim == 3 else O.spectral_conv2d(x, *ws)
    (y * dy).sum().backward()
    y_e, dx_e, dW_e = emulate_forward_backward(ndim, x.detach().numpy(), [w.detach().numpy() for w in ws],
                                               dy.numpy(), *m)
    assert O.rel_l2(torch.from_numpy(y_e), y.detach()) < 5e-7   # tables are fp32-rounded twiddles
    assert O.rel_l2(torch.from_numpy(dx_e), x.grad) < 5e-7
    for g_e, w in zip(dW_e, ws):
        assert O.rel_l2(torch.from_numpy(g_e), w.grad) < 5e-7
    if ndim == 3 and 2 * modes[0] > shape[0]:  # the overwritten part of the "low" corners really has zero gradient
        assert float(ws[0].grad.abs().min()) == 0.0
