"""CPU design study for the NEXT step of the bf16 mode (bf16 activation STORAGE) - not a test of shipped kernels.

What is built (round 2, DESIGN section 3e, GPU parity in tests/test_gpu_bf16.py) is the reference's own bf16 arithmetic:
``torch.autocast`` semantics - Linear / Conv operands in bf16 with fp32 accumulation, FFT and the complex einsum in
fp32, tensors between the kernels still fp32.  The lever that is left on the HBM-bound kernels is to additionally STORE the
activations that cross HBM between kernels (lift output, each layer's output) as bf16 - halving the `L * 2 * C * T'H'W'` term
of SURVEY 8d.  This file emulates that on the CPU oracle and checks that it stays inside the bf16 tolerance against both the
fp32 reference and the reference under autocast, i.e. that the design can meet the parity bar before any kernel is written.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import fno_oracle as O

TOL_BF16 = 1e-2  # BASELINE.json north_star


def _rb(t):
    return t.to(torch.bfloat16).to(torch.float32)


def forward_bf16_storage(ndim, sd, x, shape_out, padding=O.PADDING, training=False):
    """fno.py:105-129 (or the 2-D variant) in fp32 with the HBM-resident activations rounded to bf16."""
    L = O.n_layers_of(sd)
    if ndim == 2:
        b, t, hh, ww, ci = x.shape
        h = x.permute(0, 2, 3, 1, 4).reshape(b, hh, ww, t * ci)
        h = torch.cat((h, O.grid2d(b, hh, ww, x.dtype)), dim=-1)
        h = F.linear(h, sd["fc0.weight"], sd["fc0.bias"]).permute(0, 3, 1, 2)
        h = _rb(F.pad(h, [0, padding, 0, padding]))
    else:
        h = torch.cat((x, O.grid3d(x.shape, x.dtype)), dim=-1)
        h = F.linear(h, sd["fc0.weight"], sd["fc0.bias"]).permute(0, 4, 1, 2, 3)
        h = _rb(F.pad(h, [0, padding] * 3))
    for i in range(L):
        p = f"spectral_convs.{i}."
        if ndim == 2:
            x1 = O.spectral_conv2d(h, sd[p + "weights1"], sd[p + "weights2"])
            x2 = F.conv2d(h, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"])
        else:
            x1 = O.spectral_conv3d(h, *[sd[p + f"weights{k}"] for k in (1, 2, 3, 4)])
            x2 = F.conv3d(h, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"])
        h = O._bn(x1 + x2, sd, i, training)
        h = _rb(F.gelu(h) if i < L - 1 else h)
    if ndim == 2:
        t_out, _, _, co = shape_out
        h = h[..., :-padding, :-padding].permute(0, 2, 3, 1)
        h = F.linear(F.gelu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"])), sd["fc2.weight"], sd["fc2.bias"])
        return h.reshape(b, hh, ww, t_out, co).permute(0, 3, 1, 2, 4).contiguous()
    r = shape_out[0] // x.shape[1]
    h = h[..., :-padding, :-padding, :-padding].permute(0, 2, 3, 4, 1)
    h = F.linear(F.gelu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"])), sd["fc2.weight"], sd["fc2.bias"])
    h = h.reshape(*h.shape[:-1], shape_out[-1], r)
    return h.permute(0, 1, 5, 2, 3, 4).reshape(h.shape[0], *shape_out)


@pytest.mark.parametrize("ndim,modes,width,s", [
    (2, (16, 16), 128, (20, 64, 64, 3)),   # BASELINE C3 model (configs/fsi/fno.yaml as FNO-2D)
    (2, (12, 16), 64, (20, 64, 128, 3)),   # C2 model on the dataset's own grid
    (3, (4, 12, 16), 64, (20, 32, 64, 3)),  # the reference FNO3d (cylinder yaml) on a reduced grid
])
def test_bf16_activation_storage_is_inside_the_bf16_tolerance(ndim, modes, width, s):
    torch.manual_seed(0)
    sd = O.init_state(ndim, modes, 4, width, s, s)
    O.randomize_bn(sd)
    x = torch.randn(2, *s)
    fwd = O.fno2d_forward if ndim == 2 else O.fno3d_forward
    ref = fwd(sd, x, s)
    with torch.autocast("cpu", dtype=torch.bfloat16):  # the reference's bf16 path
        ref_autocast = fwd(sd, x, s)
    assert ref_autocast.dtype == torch.bfloat16
    got = forward_bf16_storage(ndim, sd, x, s)
    e_fp32, e_ac = O.rel_l2(got, ref), O.rel_l2(got, ref_autocast.float())
    e_ref = O.rel_l2(ref_autocast.float(), ref)
    print(f"ndim {ndim} width {width}: bf16-storage vs fp32 {e_fp32:.2e}, vs autocast {e_ac:.2e}; autocast vs fp32 {e_ref:.2e}")
    assert e_fp32 < TOL_BF16 and e_ac < TOL_BF16
    assert e_fp32 < e_ref  # closer to the fp32 result than the reference's own bf16 path is


def test_bf16_activation_storage_over_a_rollout():
    """20 autoregressive steps (BASELINE C2's rollout length) with the re-normalisation of eval.py:313-319: the bf16
    rounding is re-fed every step; the end-to-end drift against the fp32 rollout stays inside the bf16 tolerance."""
    torch.manual_seed(1)
    s = (4, 32, 48, 3)
    sd = O.init_state(2, (8, 8), 4, 32, s, s)
    O.randomize_bn(sd)
    norm = O.synthetic_normalizer(3, 3)
    x, tgt = torch.randn(2, *s), torch.randn(2, 20 * s[0], *s[1:])
    ref = O.rollout(lambda t: O.fno2d_forward(sd, t, s), norm, x, tgt, 20)[0]
    got = O.rollout(lambda t: forward_bf16_storage(2, sd, t, s), norm, x, tgt, 20)[0]
    e = O.rel_l2(got, ref)
    print(f"20-step rollout, bf16 activation storage vs fp32: {e:.2e}")
    assert e < TOL_BF16


def _param_grads(fn, sd, x, tgt):
    """train.py:328-329: loss = model.train_loss(input, target).mean(); loss.backward() on a restated forward."""
    ps = {k: v.clone().requires_grad_(True) for k, v in sd.items() if O.is_param(k)}
    full = {k: v.clone() for k, v in sd.items()}
    full.update(ps)
    F.mse_loss(fn(full, x).float(), tgt, reduction="none").mean().backward()
    real = lambda t: torch.view_as_real(t) if t.is_complex() else t
    return {k: real(p.grad) for k, p in ps.items()}


def test_bf16_activation_storage_training_gradients():
    """The C3 training step (FNO-2D fsi grid, train-mode BatchNorm): with activations AND activation gradients stored
    as bf16 between kernels (the cast's backward rounds the gradient too) every parameter gradient stays within 1e-2 of
    the fp32 step and is closer to it than the reference's own autocast step is."""
    torch.manual_seed(0)
    s = (20, 64, 64, 3)
    sd = O.init_state(2, (16, 16), 4, 64, s, s)
    O.randomize_bn(sd)
    x, tgt = torch.randn(4, *s), torch.randn(4, *s)

    def autocast_fwd(sd_, x_):
        with torch.autocast("cpu", dtype=torch.bfloat16):
            return O.fno2d_forward(sd_, x_, s, training=True)

    g_fp32 = _param_grads(lambda sd_, x_: O.fno2d_forward(sd_, x_, s, training=True), sd, x, tgt)
    g_ac = _param_grads(autocast_fwd, sd, x, tgt)
    g_bf = _param_grads(lambda sd_, x_: forward_bf16_storage(2, sd_, x_, s, training=True), sd, x, tgt)
    worst = 0.0
    for k, g in g_fp32.items():
        if k.startswith("convs.") and k.endswith(".bias"):
            continue  # analytically zero under batch-statistics BatchNorm (a per-channel shift cancels): pure rounding noise
        e_bf, e_ac = O.rel_l2(g_bf[k], g), O.rel_l2(g_ac[k], g)
        worst = max(worst, e_bf)
        assert e_bf < TOL_BF16, (k, e_bf)
        assert e_bf < e_ac, (k, e_bf, e_ac)
    print(f"worst parameter-gradient error with bf16 storage: {worst:.2e}")


# ---------------------------------------------------------------------------------------------------------------
# Which tensor-core arithmetic does the bf16 mode need?  The fp32 path pays 3 MMAs per product (3xTF32) to reach 1e-5.
# ---------------------------------------------------------------------------------------------------------------
def tf32(t):
    """what tcgen05.mma kind::tf32 does to an fp32 operand: truncate to 10 mantissa bits (profiles/r01_tf32_conversion_mode.json)"""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)

def mm(a, b, q):      # a [..., K] x b [K, N] with operand quantiser q
    return q(a) @ q(b)

def cmm(ar, ai, br, bi, q):   # complex (ar + i ai) @ (br + i bi), 4 real GEMMs
    return mm(ar, br, q) - mm(ai, bi, q), mm(ar, bi, q) + mm(ai, br, q)

def fno2d_forward_tc(sd, x, shape_out, q, store, padding=6):
    """FNO-2D forward with every tensor-core GEMM written as a matmul whose operands go through q
    (identity = fp32 / 3xTF32, tf32 = single-pass TF32) and HBM-resident activations through store."""
    b, t, hh, ww, ci = x.shape
    t_out, _, _, co = shape_out
    h = x.permute(0, 2, 3, 1, 4).reshape(b, hh, ww, t * ci)
    h = torch.cat((h, O.grid2d(b, hh, ww, x.dtype), torch.ones(b, hh, ww, 1)), dim=-1)
    w0 = torch.cat((sd["fc0.weight"], sd["fc0.bias"][:, None]), dim=1)           # bias as a K column (the lift kernel)
    h = mm(h, w0.t(), q)                                                          # [B,H,W,C] channels-last
    h = store(F.pad(h, [0, 0, 0, padding, 0, padding]))
    Hp, Wp = hh + padding, ww + padding
    L = O.n_layers_of(sd)
    m2, m3 = sd["spectral_convs.0.weights1"].shape[2:]
    FW = torch.fft.rfft(torch.eye(Wp), dim=-1)[:, :m3]                           # [Wp, m3] forward-W table
    FH = torch.fft.fft(torch.eye(Hp), dim=-1)                                     # [Hp, Hp]
    rows = torch.cat((torch.arange(m2), torch.arange(Hp - m2, Hp)))               # kept H frequencies (low, high)
    FHk = FH[:, rows]                                                             # [Hp, 2 m2]
    IH = torch.fft.ifft(torch.eye(Hp), dim=-1)[rows, :]                           # [2 m2, Hp]
    spec = torch.zeros(m3, Wp // 2 + 1, dtype=torch.cfloat); spec[torch.arange(m3), torch.arange(m3)] = 1
    GWr = torch.fft.irfft(spec, n=Wp, dim=-1)                                     # [m3, Wp]: response to Re part
    GWi = torch.fft.irfft(1j * spec, n=Wp, dim=-1)                                # [m3, Wp]: response to Im part
    for i in range(L):
        p = f"spectral_convs.{i}."
        # forward W (tc_fwdw): contraction over w
        hw = h.permute(0, 1, 3, 2)                                                # [B,H,C,W]
        ar, ai = mm(hw, FW.real, q), mm(hw, FW.imag, q)                           # [B,H,C,m3]
        # forward H (tc_tmul): contraction over h
        ar, ai = ar.permute(0, 2, 3, 1), ai.permute(0, 2, 3, 1)                   # [B,C,m3,H]
        sr, si = cmm(ar, ai, FHk.real, FHk.imag, q)                               # [B,C,m3,2m2]
        S = torch.complex(sr, si)
        Wc = torch.cat((sd[p + "weights1"], sd[p + "weights2"]), dim=2)           # [Ci,Co,2m2,m3]
        Oc = torch.einsum("bizy,ioyz->bozy", S, Wc)                               # fp32 FFMA mode mixing
        if 2 * m2 > Hp:
            raise NotImplementedError("overlapping corners not needed for this study")
        # inverse H (FFMA lmul): contraction over kept kh
        dr, di = cmm(Oc.real, Oc.imag, IH.real, IH.imag, lambda t_: t_)           # [B,Co,m3,H]
        # layer kernel: bypass conv + inverse W, BN, GELU
        dr, di = dr.permute(0, 3, 1, 2), di.permute(0, 3, 1, 2)                   # [B,H,Co,m3]
        x1 = mm(dr, GWr, q) + mm(di, GWi, q)                                      # [B,H,Co,W]
        x1 = x1.permute(0, 1, 3, 2)
        x2 = mm(h, sd[f"convs.{i}.weight"][:, :, 0, 0].t(), q) + sd[f"convs.{i}.bias"]
        z = x1 + x2
        pbn = f"bns.{i}."
        z = (z - sd[pbn + "running_mean"]) / torch.sqrt(sd[pbn + "running_var"] + 1e-5) * sd[pbn + "weight"] + sd[pbn + "bias"]
        h = store(F.gelu(z) if i < L - 1 else z)
    h = h[:, :hh, :ww]
    h = F.gelu(mm(h, sd["fc1.weight"].t(), q) + sd["fc1.bias"])
    h = mm(h, sd["fc2.weight"].t(), q) + sd["fc2.bias"]
    return h.reshape(b, hh, ww, t_out, co).permute(0, 3, 1, 2, 4).contiguous()


@pytest.mark.parametrize("modes,width,s", [((16, 16), 128, (20, 64, 64, 3)), ((12, 16), 64, (20, 64, 128, 3))])
def test_tensor_core_arithmetic_options_for_the_bf16_mode(modes, width, s):
    """FNO-2D forward with every tensor-core GEMM of the engine (lift, forward-W, forward-H, bypass conv + inverse-W,
    fc1, fc2) written as a matmul whose operands pass through a quantiser: single-pass TF32 (operands truncated, what
    `tcgen05.mma kind::tf32` does) or bf16 operands (`kind::f16`), fp32 accumulation; mode mixing and inverse-H in fp32.
    All options stay far inside the 1e-2 bf16 tolerance, so the bf16 mode can drop the 3xTF32 lo planes (a third of the
    MMAs, half the D / weight / table traffic and TMEM A-operand columns) and even run the GEMMs in bf16."""
    ident = lambda t: t
    torch.manual_seed(0)
    sd = O.init_state(2, modes, 4, width, s, s)
    O.randomize_bn(sd)
    x = torch.randn(2, *s)
    ref = O.fno2d_forward(sd, x, s)
    e_form = O.rel_l2(fno2d_forward_tc(sd, x, s, ident, ident), ref)
    e_tf32 = O.rel_l2(fno2d_forward_tc(sd, x, s, tf32, ident), ref)
    e_tf32_bf = O.rel_l2(fno2d_forward_tc(sd, x, s, tf32, _rb), ref)
    e_bf_bf = O.rel_l2(fno2d_forward_tc(sd, x, s, _rb, _rb), ref)
    print(f"width {width}: matmul form {e_form:.1e} | 1xTF32 {e_tf32:.1e} | 1xTF32 + bf16 storage {e_tf32_bf:.1e} | "
          f"bf16 operands + bf16 storage {e_bf_bf:.1e}")
    assert e_form < 1e-6                      # the matmul formulation is the reference arithmetic
    assert 1e-5 < e_tf32 < TOL_BF16           # single-pass TF32 misses the fp32 bar (why the fp32 path is 3xTF32) ...
    assert e_tf32_bf < TOL_BF16 and e_bf_bf < TOL_BF16   # ... and every bf16-mode option meets the bf16 bar
