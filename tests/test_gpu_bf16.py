"""bf16 compute mode (north_star: "1e-2 (bf16)"; BASELINE config C3 names bf16) on the GPU.

The reference has exactly one working bf16 path, ``torch.autocast(dtype=torch.bfloat16)`` around the unmodified module
(SURVEY F7): every ``nn.Linear`` / ``nn.Conv3d`` casts its operands to bf16 and accumulates in fp32, ``rfftn`` /
``irfftn`` / the complex ``einsum`` stay fp32 (fno.py:48,63,41-43), BatchNorm and GELU run in fp32 on the fp32 sum.
``B200FNO_COMPUTE_BF16`` is that arithmetic on the engine: operands of fc0, the 1x1 convolutions, fc1 and fc2 rounded
to bf16 (weights when packed, activations in the operand-staging warps), ONE tensor-core pass instead of the three
3xTF32 passes, everything spectral unchanged.  Checked here against

* the fp32 oracle, tolerance 1e-2 (the north_star bar) - measured ~2e-3;
* the oracle run under ``torch.autocast("cpu", dtype=torch.bfloat16)`` - the reference's own bf16 numbers; the two
  differ by the roundings autocast additionally applies to every Linear / Conv OUTPUT (which the fused kernels keep
  in fp32), so the bound is the autocast path's own distance from fp32 (~4e-3).
"""
import pytest
import torch

from oracle import fno_oracle as O

pytestmark = pytest.mark.gpu

TOL_BF16 = 1e-2        # north_star
TOL_VS_AUTOCAST = 6e-3


@pytest.fixture(scope="module")
def R():
    import realpdebench_b200 as R
    from realpdebench_b200 import _capi
    _capi.lib()
    return R


def dev():
    return torch.device("cuda:0")


def build(R, ndim, modes, L, width, s_in, s_out, seed, gain=1.0):
    torch.manual_seed(seed)
    sd = O.init_state(ndim, modes, L, width, s_in, s_out)
    O.randomize_bn(sd, seed + 1)
    sd = {k: (v * gain if k.startswith("spectral_convs.") else v.clone()) for k, v in sd.items()}
    m = (R.FNO3d if ndim == 3 else R.FNO2d)(*modes, L, width, s_in, s_out)
    m.load_state_dict(sd)
    return m.to(dev()).eval(), sd


def oracle_pair(ndim, sd, x, s_out):
    fwd = O.fno3d_forward if ndim == 3 else O.fno2d_forward
    with torch.no_grad():
        y32 = fwd(sd, x, s_out)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yac = fwd(sd, x, s_out)
    return y32, yac.float()


CASES = [
    (2, (6, 8), 3, 64, (4, 40, 100, 3), "tc"),      # tcgen05 path, one partial 128-point tile
    (2, (12, 16), 4, 64, (20, 64, 250, 3), "tc"),   # C2 channels / modes on a smaller grid, two full tiles
    (3, (2, 4, 8), 3, 64, (6, 20, 58, 3), "tc"),    # FNO3d, the reference's own module shape family
    (2, (5, 6), 2, 32, (4, 20, 28, 3), "simt"),     # FFMA path (width 32)
    (2, (8, 8), 2, 128, (5, 24, 24, 3), "simt"),    # fsi width (C3 model family), FFMA path
]


@pytest.mark.parametrize("gain", [1.0, 30.0])
@pytest.mark.parametrize("ndim,modes,L,width,s,impl", CASES)
def test_bf16_forward_vs_fp32_oracle_and_autocast_oracle(R, ndim, modes, L, width, s, impl, gain):
    m, sd = build(R, ndim, modes, L, width, s, s, seed=61, gain=gain)
    torch.manual_seed(4)
    x = torch.randn(2, *s)
    y32, yac = oracle_pair(ndim, sd, x, s)
    m.set_compute("bf16")
    y = m(x.to(dev())).cpu()
    assert y.dtype == torch.float32 and m.engine.resolved_impl() == impl
    e32, eac = O.rel_l2(y, y32), O.rel_l2(y, yac)
    assert 1e-5 < e32 < TOL_BF16, e32          # really reduced precision, and inside the bar
    assert eac < TOL_VS_AUTOCAST, eac
    assert e32 <= 1.5 * O.rel_l2(yac, y32)      # no worse than the reference's own autocast path
    # back to fp32: the packed weights are refreshed and 1e-5 parity returns
    m.set_compute("f32")
    assert O.rel_l2(m(x.to(dev())).cpu(), y32) < 1e-5


def test_autocast_context_selects_the_bf16_mode_and_returns_bf16(R):
    """Drop-in: the caller wraps the module in torch.autocast exactly as with the reference."""
    s = (4, 40, 100, 3)
    m, sd = build(R, 2, (6, 8), 3, 64, s, s, seed=62)
    torch.manual_seed(5)
    x = torch.randn(2, *s)
    y32, yac = oracle_pair(2, sd, x, s)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x.to(dev()))
    assert y.dtype == torch.bfloat16 and m.engine.compute == "bf16"
    assert O.rel_l2(y.float().cpu(), y32) < TOL_BF16
    assert O.rel_l2(y.float().cpu(), yac) < TOL_VS_AUTOCAST + 4e-3  # + the final bf16 rounding of both outputs
    y2 = m(x.to(dev()))  # outside the context: fp32 again
    assert y2.dtype == torch.float32 and O.rel_l2(y2.cpu(), y32) < 1e-5


def test_bf16_rollout_3_steps(R):
    """The fused rollout in bf16 mode against the fp32 oracle rollout: per-step slices within the bf16 bar."""
    s = (4, 40, 100, 3)
    m, sd = build(R, 2, (6, 8), 3, 64, s, s, seed=63)
    norm = O.synthetic_normalizer(3, 3, seed=99)
    torch.manual_seed(6)
    n = 3
    x, tgt = torch.randn(2, *s), torch.randn(2, n * 4, 40, 100, 3)
    with torch.no_grad():
        pred_o, _, loss_o, _ = O.rollout(lambda t: O.fno2d_forward(sd, t, s), norm, x, tgt, n)
    import copy
    nd = copy.copy(norm)
    nd.device = dev()
    for k, v in vars(norm).items():
        if torch.is_tensor(v):
            setattr(nd, k, v.to(dev()))
    m.set_compute("bf16")
    pred, _, loss, _ = R.rollout(m, nd, x.to(dev()), tgt.to(dev()), n)
    for i in range(n):
        sl = slice(4 * i, 4 * i + 4)
        assert O.rel_l2(pred[:, sl].cpu(), pred_o[:, sl]) < TOL_BF16, f"step {i}"
    assert abs(loss - loss_o) < 1e-2 * max(1.0, abs(loss_o))


TRAIN_CASES = [
    (2, (5, 6, 2, 12, (4, 20, 28, 3), (4, 20, 28, 3)), 3),
    (3, (2, 3, 3, 2, 8, (3, 7, 9, 2), (3, 7, 9, 2)), 2),
    (2, (4, 5, 2, 128, (2, 9, 10, 2), (2, 9, 10, 2)), 6),     # width 128: the C3 model family
]


@pytest.mark.parametrize("ndim,ctor,batch", TRAIN_CASES)
@pytest.mark.parametrize("how", ["set_compute", "autocast"])
def test_bf16_training_step_vs_autocast_oracle(R, ndim, ctor, batch, how):
    """BASELINE config C3 names bf16 training.  In bf16 mode both operands of every Linear / Conv GEMM - forward, grad_input
    and grad_weight - are bf16 tensors with fp32 accumulation, which is what autograd does under
    ``torch.autocast(bfloat16)``; the spectral stages, BatchNorm and the loss stay fp32.  Checked against autograd through
    the oracle under CPU autocast (the reference's own bf16 numbers, themselves 4e-3 .. 1e-2 from fp32) and against the
    fp32 oracle (the north_star's 1e-2 bar, per gradient tensor)."""
    torch.manual_seed(14)
    m = (R.FNO3d if ndim == 3 else R.FNO2d)(*ctor)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    O.randomize_bn(sd, 23)
    m.load_state_dict(sd)
    m = m.to(dev()).train()
    s_in, s_out = ctor[-2], ctor[-1]
    x, t = torch.randn(batch, *s_in), torch.randn(batch, *s_out)
    l32, g32, _ = O.train_loss_and_grads(ndim, {k: v.clone() for k, v in sd.items()}, x, t, s_out, input_grad=True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lac, gac, _ = O.train_loss_and_grads(ndim, {k: v.clone() for k, v in sd.items()}, x, t, s_out, input_grad=True)
    xd = x.to(dev()).requires_grad_(True)
    if how == "autocast":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = m.train_loss(xd, t.to(dev())).mean()
    else:
        m.set_compute("bf16")
        loss = m.train_loss(xd, t.to(dev())).mean()
    loss.backward()
    assert m.engine.compute == "bf16"
    assert abs(loss.item() - l32) < 3e-3 * abs(l32)
    got = {k: p.grad.cpu() for k, p in m.named_parameters()}
    got["__input__"] = xd.grad.cpu()
    worst32 = worst_ac = 0.0
    for k, g in got.items():
        if k.startswith("convs.") and k.endswith(".bias"):
            continue  # exact gradient is zero (train-mode BatchNorm removes the mean): rounding noise only
        e32, eac = O.rel_l2(g, g32[k]), O.rel_l2(g, gac[k].to(g.dtype))
        worst32, worst_ac = max(worst32, e32), max(worst_ac, eac)
        assert e32 < 1.5e-2 and eac < 2e-2, (k, e32, eac)
    assert worst32 > 1e-4, worst32  # the mode really is reduced precision (fp32 mode: ~1e-6)
    m.set_compute("f32")
