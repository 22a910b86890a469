"""GPU parity for SURVEY 8f row N4: surrogate materialisation and the sibling spectral layers on the engine, against
the golden outputs recorded from the reference (tests/golden/surrogate.pt, siblings.pt) and the CPU oracle.
Tolerance: relative L2 <= 1e-5 (fp32, BASELINE.json north_star).  The surrogate output sits on a large constant
(``mean_targets`` plus the network's output bias), so the fluctuation around the mean is checked as well, at 2e-4: a
wrong network output shows up there at O(1)."""
import numpy as np
import pytest
import torch

from oracle import fno_oracle as O
from oracle import widening_oracle as WO
from test_widening_cpu import centred_rel_l2, sparseKernelFT2d, surrogate_traj

pytestmark = pytest.mark.gpu

TOL = 1e-5
TOL_CENTRED = 2e-4


def dev():
    return torch.device("cuda:0")


def engine_model(sd, ctor):
    import realpdebench_b200 as R
    from realpdebench_b200 import _capi
    _capi.lib()  # fail loudly if the CUDA library is missing
    m = R.FNO3d(*ctor)
    m.load_state_dict(sd)
    return m.to(dev()).eval()


# ---------------------------------------------------------------- surrogate materialisation
def test_surrogate_matches_reference_script_golden(golden):
    from realpdebench_b200.surrogate import materialize_surrogate
    g = golden("surrogate.pt")
    m = engine_model(g["sd"], g["ctor"])
    norm = O.Normalizer("gaussian", **g["norm"])
    traj = surrogate_traj(g["seed"], g["n"])
    got = materialize_surrogate(m, norm, traj, g["gas_ratio"], g["equivalence_ratio"], step=g["step"],
                                batch_size=g["batch_size"], sub_s=g["sub_s"])
    want = g["pred_traj"]
    assert got.shape == tuple(want.shape) and got.dtype == np.float32
    assert O.rel_l2(torch.from_numpy(got), want) < TOL
    assert centred_rel_l2(got, want) < TOL_CENTRED


def test_surrogate_script_model_config_vs_oracle():
    """The script's own model (generate_surrogate_data.py:27-35: modes (4,16,16), 4 layers, width 64, 10-frame windows of
    128 x 128 x 17 -> 1 channel): tensor-core kernels, C_in != C_out, single output channel."""
    from realpdebench_b200.surrogate import materialize_surrogate
    torch.manual_seed(100)
    s_in, s_out = (10, 128, 128, 17), (10, 128, 128, 1)
    sd = O.init_state(3, (4, 16, 16), 4, 64, s_in, s_out)
    O.randomize_bn(sd, 101)
    norm = O.synthetic_normalizer(17, 1, seed=102)
    traj = torch.randn(11, 128, 128, 15).numpy()
    want = WO.materialize_surrogate(lambda x: O.fno3d_forward(sd, x, s_out), norm, traj, 40, 0.85, 10, 1)
    m = engine_model(sd, (4, 16, 16, 4, 64, s_in, s_out))
    got = materialize_surrogate(m, norm, traj, 40, 0.85, step=10, batch_size=1)
    assert m.engine.resolved_impl() == "tc"
    assert got.shape == want.shape == (11, 128, 128)
    assert O.rel_l2(torch.from_numpy(got), torch.from_numpy(want)) < TOL
    assert centred_rel_l2(got, want) < TOL_CENTRED


def test_surrogate_subsampled_range_normaliser_vs_oracle():
    from realpdebench_b200.surrogate import materialize_surrogate
    torch.manual_seed(90)
    step, c = 3, 4
    s_in, s_out = (step, 6, 5, c + 2), (step, 6, 5, 1)  # 12 x 10 frames, sub_s = 2
    sd = O.init_state(3, (2, 2, 2), 2, 6, s_in, s_out)
    O.randomize_bn(sd, 91)
    norm = O.synthetic_normalizer(c + 2, 1, seed=92, kind="range")
    traj = torch.randn(13, 12, 10, c).double().numpy()
    want = WO.materialize_surrogate(lambda x: O.fno3d_forward(sd, x, s_out), norm, traj, 60, 1.1, step, 2, 2)
    got = materialize_surrogate(engine_model(sd, (2, 2, 2, 2, 6, s_in, s_out)), norm, traj, 60, 1.1, step=step,
                                batch_size=2, sub_s=2)
    assert O.rel_l2(torch.from_numpy(got), torch.from_numpy(want)) < TOL


# ---------------------------------------------------------------- sibling spectral layers
class sparseKernelFT3d(torch.nn.Module):
    """Stand-in with the reference class name / attributes (MWT_libs/models.py:535-555)."""

    def __init__(self, sd, modes):
        super().__init__()
        self.modes = modes
        for k in (1, 2, 3, 4):
            setattr(self, f"weights{k}", torch.nn.Parameter(sd[f"weights{k}"].clone()))
        self.Lo = torch.nn.Linear(sd["Lo.weight"].shape[1], sd["Lo.weight"].shape[0])
        with torch.no_grad():
            self.Lo.weight.copy_(sd["Lo.weight"]), self.Lo.bias.copy_(sd["Lo.bias"])
        self.reference_calls = 0

    def forward(self, x):
        self.reference_calls += 1
        ws = [self.weights1, self.weights2, self.weights3, self.weights4]
        return WO.mwt_sparse_kernel_ft3d(x, ws, self.modes, self.Lo.weight, self.Lo.bias)


class SpectralConv3d(torch.nn.Module):
    """Stand-in for galerkin_transformer_libs/layers.py:1205-1257."""

    def __init__(self, sd):
        super().__init__()
        for k in (1, 2, 3, 4):
            setattr(self, f"weights{k}", torch.nn.Parameter(sd[f"weights{k}"].clone()))
        self.reference_calls = 0

    def forward(self, x):
        self.reference_calls += 1
        return WO.galerkin_spectral_conv3d(x, [self.weights1, self.weights2, self.weights3, self.weights4])


@pytest.mark.parametrize("case", ["ft3d", "ft3d_clipped", "ft2d", "ft2d_clipped", "galerkin3d"])
def test_routed_sibling_layers_match_reference_golden(golden, case):
    from realpdebench_b200 import _capi, siblings
    _capi.lib()
    g = golden("siblings.pt")[case]
    if g["kind"] == "ft3d":
        m = sparseKernelFT3d(g["sd"], g["alpha"])
    elif g["kind"] == "ft2d":
        m = sparseKernelFT2d(g["sd"], g["alpha"])
    else:
        m = SpectralConv3d(g["sd"])
    m = m.to(dev()).eval()
    assert siblings.route(m)
    with torch.no_grad():
        y = m(g["x"].to(dev()))
    assert m.reference_calls == 0  # served by b200fno_spectral_conv, not by the module's torch forward
    assert y.shape == g["y"].shape
    assert O.rel_l2(y.cpu(), g["y"]) < TOL
    if case == "ft2d":  # with autograd on, the call is the reference forward (the engine operator has no backward)
        m(g["x"].to(dev())).sum().backward()
        assert m.reference_calls == 1 and m.weights1.grad is not None


# ---------------------------------------------------------------- edge case: empty batch
def test_empty_batch_is_an_error_like_the_reference(golden):
    """The reference fails on a zero-sample batch (its FFT rejects the empty transform with a RuntimeError, checked on
    the CPU build: 'MKL FFT error ... Inconsistent configuration parameters'); the engine reports a RuntimeError too
    (C-ABI: batch outside [1, max_batch]) and keeps serving real batches afterwards."""
    g = golden("kat_a.pt")
    m = engine_model(g["sd"], g["ctor"])
    x = g["x"].to(dev())
    y = m(x)
    with pytest.raises(RuntimeError):
        m(x[:0])
    m2 = engine_model(g["sd"], g["ctor"])  # no plan yet: the descriptor itself is rejected
    with pytest.raises(RuntimeError):
        m2(x[:0])
    assert torch.equal(m(x), y)
