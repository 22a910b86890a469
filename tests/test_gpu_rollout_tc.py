"""GPU parity of the HEADLINE path: the width-64 tensor-core kernels over a multi-step rollout.

Every BENCH / SCALE number after the first autoregressive step is produced by ``tc_proj_kernel`` writing the
fed-back model input (``tc_proj.cu``: the TMA store through ``tmState`` when ``c_in == c_out``, the scattered
``a.state[...]`` store when parameter channels interleave) and the tcgen05 lift reading it back.  The golden
rollouts of ``test_gpu_parity.py`` are width 8 (FFMA projection), so they never touch that code; these tests do:

* 3-step rollouts on the tcgen05 path - plain (Gaussian), Range, controlled (``c_in = c_out + 2``) - against
  ``oracle.fno_oracle.rollout`` (= eval.py:296-326): free-running <= 5e-5, teacher-forced per step <= 1e-5, loss;
* the same with the spectral weights amplified: with the reference initialisation (``scale = 1/(Ci*Co)``,
  fno.py:27) the spectral branch contributes < 1e-3 of the output, so a whole-network tolerance of 1e-5 would let a
  1 % error of the truncated-DFT stages through; a gain of 300 makes both branches comparable;
* FNO3d (time unfold r = 1) at width 64;
* the C2 benchmark shape, batch 2, all 20 steps, against the oracle functions evaluated in fp64 on the GPU;
* a negative control: the oracle fed a corrupted state differs by far more than the tolerance, i.e. the
  comparison is sensitive to what is fed back.
"""
import pytest
import torch

from oracle import fno_oracle as O

pytestmark = pytest.mark.gpu

TOL_STEP, TOL_FREE = 1e-5, 5e-5


@pytest.fixture(scope="module")
def R():
    import realpdebench_b200 as R
    from realpdebench_b200 import _capi
    _capi.lib()
    return R


def dev():
    return torch.device("cuda:0")


def amplified(sd, gain):
    return {k: (v * gain if k.startswith("spectral_convs.") else v.clone()) for k, v in sd.items()}


def to64(sd, device):
    out = {}
    for k, v in sd.items():
        v = v.to(device)
        out[k] = v.double() if v.is_floating_point() else (v.to(torch.cdouble) if v.is_complex() else v)
    return out


def norm_on(norm, device, dtype=None):
    import copy
    n = copy.copy(norm)
    n.device = device
    for k, v in vars(norm).items():
        if torch.is_tensor(v):
            setattr(n, k, v.to(device=device, dtype=dtype or v.dtype))
    return n


def build(R, ndim, modes, L, s_in, s_out, seed, gain=1.0):
    torch.manual_seed(seed)
    sd = O.init_state(ndim, modes, L, 64, s_in, s_out)
    O.randomize_bn(sd, seed + 1)
    sd = amplified(sd, gain)
    m = R.FNO3d(*modes, L, 64, s_in, s_out) if ndim == 3 else R.FNO2d(*modes, L, 64, s_in, s_out)
    m.load_state_dict(sd)
    return m.to(dev()).eval(), sd


def assert_tc_path(m, n_auto):
    si = m.engine.stage_impls()
    for s in ("lift", "fwdW", "modes", "layer", "proj"):
        assert si[s] == "tc", (s, si)


@pytest.mark.parametrize("gain", [1.0, 300.0])
@pytest.mark.parametrize("case", ["plain", "range", "controlled"])
def test_tc_rollout_3step_fno2d_width64(R, case, gain):
    c_out = 3
    c_in = 5 if case == "controlled" else 3
    s_in, s_out = (4, 40, 100, c_in), (4, 40, 100, c_out)
    m, sd = build(R, 2, (6, 8), 3, s_in, s_out, seed=21, gain=gain)
    norm = O.synthetic_normalizer(c_in, c_out, seed=99, kind="range" if case == "range" else "gaussian")
    torch.manual_seed(5)
    n = 3
    x = torch.randn(2, *s_in)
    tgt = torch.randn(2, n * 4, 40, 100, c_out)
    fwd = lambda t: O.fno2d_forward(sd, t, s_out)
    with torch.no_grad():
        pred_o, tgt_o, loss_o, states = O.rollout(fwd, norm, x, tgt, n)
    pred, tgt_e, loss, _ = R.rollout(m, norm_on(norm, dev()), x.to(dev()), tgt.to(dev()), n)
    assert_tc_path(m, n)
    assert pred.shape == pred_o.shape
    assert O.rel_l2(pred.cpu(), pred_o) < TOL_FREE
    for i in range(n):  # every step's slice on its own: a wrong fed-back state shows from slice 1 on
        sl = slice(4 * i, 4 * i + 4)
        assert O.rel_l2(pred[:, sl].cpu(), pred_o[:, sl]) < TOL_FREE, f"step {i}"
    assert O.rel_l2(tgt_e.cpu(), tgt_o) < 1e-6
    assert abs(loss - loss_o) < 1e-4 * max(1.0, abs(loss_o))
    # teacher forced: each step from the oracle's own state; multi-step engine call vs single steps
    a, b = R.rollout_affine(norm_on(norm, dev()), c_in, c_out, dev())
    for i in range(n):
        step = m.rollout(states[i].to(dev()), a, b, 1).cpu()
        assert O.rel_l2(step, states[i + 1][..., :c_out]) < TOL_STEP, f"teacher-forced step {i}"
    # negative control: the oracle fed a corrupted state is far outside the tolerance
    with torch.no_grad():
        bad = fwd(states[1].roll(1, dims=1))
        good = fwd(states[1])
    assert O.rel_l2(bad, good) > 20 * TOL_FREE


@pytest.mark.parametrize("gain", [1.0, 300.0])
def test_tc_rollout_chained_equals_single_steps(R, gain):
    """The n-step engine call (state ping-pong inside b200fno_rollout) == n one-step calls fed by hand, bit for bit."""
    s = (4, 40, 100, 3)
    m, _ = build(R, 2, (6, 8), 3, s, s, seed=23, gain=gain)
    norm = norm_on(O.synthetic_normalizer(3, 3, seed=7), dev())
    a, b = R.rollout_affine(norm, 3, 3, dev())
    torch.manual_seed(8)
    x0 = torch.randn(3, *s, device=dev())
    chained = m.rollout(x0, a, b, 4)
    cur = x0
    for i in range(4):
        cur = m.rollout(cur, a, b, 1)
        assert torch.equal(cur, chained[:, 4 * i:4 * i + 4]), f"step {i}"


@pytest.mark.parametrize("gain", [1.0, 300.0])
def test_tc_rollout_3step_fno3d_width64_r1(R, gain):
    s = (6, 20, 58, 3)  # padded (12, 26, 64): one 64-point tile per row
    m, sd = build(R, 3, (2, 4, 8), 3, s, s, seed=31, gain=gain)
    norm = O.synthetic_normalizer(3, 3, seed=17)
    torch.manual_seed(6)
    n = 3
    x = torch.randn(2, *s)
    tgt = torch.randn(2, n * 6, 20, 58, 3)
    fwd = lambda t: O.fno3d_forward(sd, t, s)
    with torch.no_grad():
        pred_o, _, loss_o, states = O.rollout(fwd, norm, x, tgt, n)
    pred, _, loss, _ = R.rollout(m, norm_on(norm, dev()), x.to(dev()), tgt.to(dev()), n)
    assert_tc_path(m, n)
    assert O.rel_l2(pred.cpu(), pred_o) < TOL_FREE
    assert abs(loss - loss_o) < 1e-4 * max(1.0, abs(loss_o))
    a, b = R.rollout_affine(norm_on(norm, dev()), 3, 3, dev())
    for i in range(n):
        step = m.rollout(states[i].to(dev()), a, b, 1).cpu()
        assert O.rel_l2(step, states[i + 1]) < TOL_STEP, f"teacher-forced step {i}"


@pytest.mark.parametrize("gain", [1.0, 300.0])
def test_c2_shape_rollout20_vs_fp64_oracle_on_gpu(R, gain):
    """BASELINE config C2 (FNO-2D 256x512, 20 frames x 3 ch, modes (12,16), width 64, 4 layers), batch 2, the full
    20-step rollout of bench.py, against the oracle's functions evaluated in float64 on the GPU (same code as the
    CPU oracle, other device / precision).  Also pins bench.py's ``normalized_loss_check``."""
    s = (20, 256, 512, 3)
    m, sd = build(R, 2, (12, 16), 4, s, s, seed=41, gain=gain)
    norm = O.synthetic_normalizer(3, 3)
    torch.manual_seed(9)
    n = 20
    x = torch.randn(2, *s)
    tgt = torch.randn(2, n * 20, 256, 512, 3)
    sd64 = to64(sd, dev())
    norm64 = norm_on(norm, dev(), torch.float64)
    with torch.no_grad():
        pred_o, _, loss_o, states = O.rollout(lambda t: O.fno2d_forward(sd64, t, s), norm64, x.to(dev()).double(),
                                              tgt.to(dev()).double(), n)
    pred, _, loss, _ = R.rollout(m, norm_on(norm, dev()), x.to(dev()), tgt.to(dev()), n)
    assert_tc_path(m, n)
    assert O.rel_l2(pred, pred_o) < TOL_FREE
    for i in (0, 1, 2, 10, 19):
        sl = slice(20 * i, 20 * i + 20)
        assert O.rel_l2(pred[:, sl], pred_o[:, sl]) < TOL_FREE, f"step {i}"
    assert abs(loss - loss_o) < 1e-5 * max(1.0, abs(loss_o))
    a, b = R.rollout_affine(norm_on(norm, dev()), 3, 3, dev())
    for i in (0, 7, 19):
        step = m.rollout(states[i].float(), a, b, 1)
        assert O.rel_l2(step, states[i + 1]) < TOL_STEP, f"teacher-forced step {i}"


@pytest.mark.parametrize("ndim,modes,s", [
    (2, (12, 16), (2, 20, 250, 2)),     # two full 128-point tiles
    (3, (2, 4, 16), (4, 10, 60, 2)),    # 3-D rows, single 72-point tile
    (2, (16, 32), (1, 60, 250, 3)),     # modes3 = 32: the widest inverse-W the layer kernel's TMEM holds
])
def test_tc_forward_with_amplified_spectral_branch(R, ndim, modes, s):
    """Whole-network parity with the spectral branch as large as the bypass (see the module docstring)."""
    m, sd = build(R, ndim, modes, 3, s, s, seed=51, gain=300.0)
    torch.manual_seed(3)
    x = torch.randn(3, *s)
    with torch.no_grad():
        ref = (O.fno3d_forward if ndim == 3 else O.fno2d_forward)(sd, x, s)
        ref0 = (O.fno3d_forward if ndim == 3 else O.fno2d_forward)(amplified(sd, 0.0), x, s)
    assert O.rel_l2(ref0, ref) > 0.05  # the spectral branch really matters in this test
    y = m(x.to(dev())).cpu()
    assert m.engine.resolved_impl() == "tc"
    assert O.rel_l2(y, ref) < TOL_STEP
    m.set_impl("simt")
    assert O.rel_l2(m(x.to(dev())).cpu(), ref) < TOL_STEP
