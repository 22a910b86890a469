"""CPU study: error of one truncated-DFT spectral conv under different 3xTF32 splittings (vs fp64)."""
import numpy as np, torch, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from realpdebench_b200 import _capi

def trunc(x):
    u = x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)
def rn(x):
    u = x.astype(np.float32).view(np.uint32)
    u = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)
def split(x, mode):
    x = x.astype(np.float32)
    if mode == "trunc":
        hi = trunc(x); lo = trunc(x - hi)   # MMA truncates lo
    elif mode == "rn":
        hi = rn(x); lo = rn(x - hi)
    elif mode == "rn_hi_trunc_lo":
        hi = rn(x); lo = trunc(x - hi)
    return hi.astype(np.float64), lo.astype(np.float64)
def mm(a, b, mode, amode=None):
    """a [M,K] @ b [K,N] -> fp32-rounded result"""
    if mode == "fp64": return a.astype(np.float64) @ b.astype(np.float64)
    if mode == "fp32": return (a.astype(np.float32) @ b.astype(np.float32))
    ah, al = split(a, amode or mode); bh, bl = split(b, mode)
    return (ah @ bh + al @ bh + ah @ bl).astype(np.float32)

def spectral2d(x, w1, w2, m2, m3, mode, amode=None):
    B, Ci, H, W = x.shape; Co = w1.shape[1]
    tabs = [_capi.host_table(2, 1, H, W, 1, m2, m3, k) for k in range(6)]
    (LF, ft, fh), LH, LHi, Gt = tabs[0], tabs[1][0], tabs[4][0], tabs[5][0]
    KH = len(fh)
    dt = np.float64 if mode == "fp64" else np.float32
    act = np.transpose(x, (0, 2, 3, 1)).astype(dt)  # B,H,W,C
    # fwdW: for each (b,h): [2m3 x W] @ [W x C]
    A = mm(LF[:2*m3, :W], act.transpose(2, 0, 1, 3).reshape(W, -1), mode, amode).reshape(2*m3, B, H, Ci)  # m,b,h,c
    # fwdH: rows (h,ri) -> (ri,kh)
    A2 = A.reshape(2, m3, B, H, Ci).transpose(3, 0, 2, 1, 4).reshape(H*2, B*m3*Ci)  # (h,ri),(b,kw,c)
    Bh = mm(LH[:2*KH, :2*H], A2, mode, amode).reshape(2, KH, B, m3, Ci)
    Sc = (Bh[0].astype(np.complex128) + 1j*Bh[1].astype(np.complex128))  # kh,b,kw,c
    Oc = np.zeros((KH, B, m3, Co), dtype=np.complex128)
    for b_, f_h in enumerate(fh):
        hi = f_h >= H - m2
        y = f_h - (H - m2) if hi else f_h
        wc = (w2 if hi else w1)[:, :, y, :]
        Oc[b_] = np.einsum("bzi,ioz->bzo", Sc[b_], wc)
    if mode != "fp64":
        Oc = Oc.astype(np.complex64)
    Or = np.stack([Oc.real, Oc.imag], 0).reshape(2*KH, B*m3*Co)
    D = mm(LHi[:2*H, :2*KH], Or, mode, amode).reshape(H, 2, B, m3, Co)  # (h,ri),b,kw,c
    D2 = D.transpose(1, 3, 0, 2, 4).reshape(2*m3, H*B*Co)
    y = mm(Gt[:W, :2*m3], D2, mode, amode).reshape(W, H, B, Co)
    return y.transpose(2, 3, 1, 0)

rng = np.random.default_rng(0)
for (H, W, m2, m3, C) in [(66, 256, 16, 32, 64), (262//2, 518//2, 12, 16, 64)]:
    x = rng.standard_normal((2, C, H, W)).astype(np.float32)
    w1 = (rng.random((C, C, m2, m3)) + 1j*rng.random((C, C, m2, m3))).astype(np.complex64) / (C*C)
    w2 = (rng.random((C, C, m2, m3)) + 1j*rng.random((C, C, m2, m3))).astype(np.complex64) / (C*C)
    ref = spectral2d(x, w1, w2, m2, m3, "fp64")
    xt = torch.from_numpy(x)
    for mode, amode in [("trunc", "rn"), ("fp32", None), ("trunc", None), ("rn", None), ("rn_hi_trunc_lo", None), ("rn", "trunc"), ("rn", "rn_hi_trunc_lo")]:
        got = spectral2d(x, w1, w2, m2, m3, mode, amode)
        e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(H, W, mode, amode, f"{e:.3e}")
    # torch fft fp32
    xf = torch.fft.rfft2(xt)
    out = torch.zeros(2, C, H, W//2+1, dtype=torch.cfloat)
    out[:, :, :m2, :m3] = torch.einsum("bixy,ioxy->boxy", xf[:, :, :m2, :m3], torch.from_numpy(w1))
    out[:, :, -m2:, :m3] = torch.einsum("bixy,ioxy->boxy", xf[:, :, -m2:, :m3], torch.from_numpy(w2))
    yt = torch.fft.irfft2(out, s=(H, W)).numpy()
    print(H, W, "torch-fft fp32", f"{np.linalg.norm(yt - ref)/np.linalg.norm(ref):.3e}")
