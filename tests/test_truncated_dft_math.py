"""CPU check of the engine's algorithm: emulate the CUDA stage sequence with numpy,
using the *library's own* constant tables (b200fno_host_table, no device needed) and
the engine's intermediate layouts, and compare with the oracle's SpectralConv."""
import numpy as np
import pytest
import torch

from oracle import fno_oracle as O
from realpdebench_b200 import _capi


def emulate_spectral(ndim, x, weights, m1, m2, m3, kw0=0):
    """x: [B,Ci,(T,)H,W] float64 numpy; weights: list of complex corner arrays.  Mirrors api.cu:run_spectral.
    ``kw0``: the mode slice [kw0, kw0 + m3) of a plan that splits modes3 (api.cu: ModeSlice); ``weights`` then hold that
    slice of the W modes."""
    if ndim == 2:
        x = x[:, :, None]
    B, Ci, T, H, W = x.shape
    Co = weights[0].shape[1]
    tabs = [_capi.host_table(ndim, T, H, W, m1, m2, m3, k, kw0) for k in range(6)]
    (LF, ft, fh), LH, LT, LTi, LHi, Gt = tabs[0], tabs[1][0], tabs[2][0], tabs[3][0], tabs[4][0], tabs[5][0]
    KT, KH = len(ft), len(fh)
    act = np.transpose(x, (0, 2, 3, 4, 1))  # channels-last [B,T,H,W,C]
    # fwd W: A[b,t,h,(ri,kw),c]
    A = np.einsum("mw,bthwc->bthmc", LF[:2 * m3, :W].astype(np.float64), act)
    # fwd H: k=(h,ri) -> m=(ri,kh):  [B,T,(ri,kh),kw,c]
    A2 = A.reshape(B, T, H, 2, m3, Ci).reshape(B, T, H * 2, m3, Ci)
    Bh = np.einsum("mk,btknc->btmnc", LH[:2 * KH, :2 * H].astype(np.float64), A2)
    if ndim == 3:
        B2 = Bh.reshape(B, T * 2, KH, m3, Ci)
        S = np.einsum("mk,bkhnc->bmhnc", LT[:2 * KT, :2 * T].astype(np.float64), B2)  # [B,(ri,kt),kh,kw,c]
    else:
        S = Bh.reshape(B, 2, KH, m3, Ci)[:, :, None].reshape(B, 2 * KT, KH, m3, Ci)
    S = S.reshape(B, 2, KT, KH, m3, Ci)
    Sc = S[:, 0] + 1j * S[:, 1]
    # per-mode mixing with the corner rule of pack.cu
    Oc = np.zeros((B, KT, KH, m3, Co), dtype=np.complex128)
    for a, f_t in enumerate(ft):
        for b_, f_h in enumerate(fh):
            h_hi = f_h >= H - m2
            y = f_h - (H - m2) if h_hi else f_h
            if ndim == 3:
                t_hi = f_t >= T - m1
                xx = f_t - (T - m1) if t_hi else f_t
                wc = weights[(2 if h_hi else 0) + (1 if t_hi else 0)][:, :, xx, y, :]
            else:
                wc = weights[1 if h_hi else 0][:, :, y, :]
            Oc[:, a, b_] = np.einsum("bzi,ioz->bzo", Sc[:, a, b_], wc)
    Or = np.stack([Oc.real, Oc.imag], axis=1)  # [B,ri,KT,KH,m3,Co]
    if ndim == 3:
        Ct = np.einsum("mk,bkhnc->bmhnc", LTi[:2 * T, :2 * KT].astype(np.float64), Or.reshape(B, 2 * KT, KH, m3, Co))
        Ct = Ct.reshape(B, T, 2 * KH, m3, Co)  # m=(t,ri) -> [B,T,(ri,kh),...]
    else:
        Ct = Or.reshape(B, 1, 2 * KH, m3, Co)
    D = np.einsum("mk,btknc->btmnc", LHi[:2 * H, :2 * KH].astype(np.float64), Ct)  # m=(h,ri)
    D = D.reshape(B, T, H, 2 * m3, Co)
    y = np.einsum("wk,bthkc->bthwc", Gt[:W, :2 * m3].astype(np.float64), D)
    y = np.transpose(y, (0, 4, 1, 2, 3))
    return y[:, :, 0] if ndim == 2 else y


@pytest.mark.parametrize("shape,modes,ci,co", [
    ((9, 10, 12), (2, 3, 4), 4, 5),      # KAT-B geometry
    ((8, 7, 16), (3, 2, 9), 3, 3),       # Nyquist bin kept (m3 = W/2+1), odd H
    ((5, 6, 7), (3, 4, 4), 2, 4),        # overlapping corners on T and H (2*m > N), m3 = W//2+1 with odd W
])
def test_3d_stage_sequence_matches_oracle(shape, modes, ci, co):
    torch.manual_seed(0)
    T, H, W = shape
    m1, m2, m3 = modes
    x = torch.randn(2, ci, T, H, W, dtype=torch.float64)
    ws = [torch.randn(ci, co, m1, m2, m3, dtype=torch.cdouble) for _ in range(4)]
    ref = O.spectral_conv3d(x, *ws)
    got = emulate_spectral(3, x.numpy(), [w.numpy() for w in ws], m1, m2, m3)
    assert O.rel_l2(torch.from_numpy(got), ref) < 5e-7  # tables are fp32-rounded twiddles


@pytest.mark.parametrize("shape,modes", [((14, 18), (5, 5)), ((9, 8), (5, 5)), ((262 // 8, 518 // 8), (12, 16))])
def test_2d_stage_sequence_matches_oracle(shape, modes):
    torch.manual_seed(1)
    H, W = shape
    m2, m3 = modes
    x = torch.randn(2, 3, H, W, dtype=torch.float64)
    ws = [torch.randn(3, 4, m2, m3, dtype=torch.cdouble) for _ in range(2)]
    ref = O.spectral_conv2d(x, *ws)
    got = emulate_spectral(2, x.numpy(), [w.numpy() for w in ws], 1, m2, m3)
    assert O.rel_l2(torch.from_numpy(got), ref) < 5e-7


@pytest.mark.parametrize("ndim,shape,modes", [(2, (30, 132), (6, 48)), (2, (22, 130), (5, 64)), (3, (8, 10, 100), (2, 3, 48))])
def test_two_mode_slices_add_up_to_the_full_operator(ndim, shape, modes):
    """modes3 in (32, 64] runs as two W-mode slices (api.cu: ModeSlice): slice s uses DFT tables with the frequency offset
    kw0 = s * m3 / 2 and the weights' W modes [kw0, kw0 + m3/2); the inverse-W terms of the two slices add.  Emulated here
    with the library's own slice tables (b200fno_host_table_slice) against the oracle's full SpectralConv."""
    torch.manual_seed(4)
    m3 = modes[-1]
    ci, co = 3, 4
    x = torch.randn(2, ci, *shape, dtype=torch.float64)
    ncorner = 4 if ndim == 3 else 2
    ws = [torch.randn(ci, co, *modes, dtype=torch.cdouble) for _ in range(ncorner)]
    ref = (O.spectral_conv3d if ndim == 3 else O.spectral_conv2d)(x, *ws)
    m1, m2 = (modes[0], modes[1]) if ndim == 3 else (1, modes[0])
    half = m3 // 2
    got = sum(emulate_spectral(ndim, x.numpy(), [w[..., s * half:(s + 1) * half].numpy() for w in ws], m1, m2, half,
                               kw0=s * half) for s in range(2))
    assert O.rel_l2(torch.from_numpy(got), ref) < 5e-7
