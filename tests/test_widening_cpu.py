"""CPU tests for SURVEY 8f row N4: the surrogate materialisation caller and the sibling spectral layers.

* the oracle restatements (oracle/widening_oracle.py) against outputs of the reference itself
  (tests/golden/surrogate.pt, siblings.pt written by ``make_golden.py widening``)
* the host logic of ``realpdebench_b200.surrogate`` (window plan, affine fold, staging / ordering) with the device
  plumbing replaced by host stand-ins and the engine call replaced by the oracle
* routing rules of ``realpdebench_b200.siblings`` and the loud CPU failure of the engine operator
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import fno_oracle as O
from oracle import widening_oracle as WO

TOL = 2e-6


def surrogate_traj(seed, n):
    """Same recipe as tests/golden/make_golden.py:surrogate_inputs (the 8.8 MB input is regenerated, not stored)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, n).reshape(n, 1, 1, 1)
    return (torch.randn(n, 128, 128, 15, generator=g) * 0.5 + torch.sin(6.0 * t)).numpy()


def centred_rel_l2(a, b):
    """rel-L2 of the fluctuation: the surrogate output sits on a large constant (mean_targets), which would hide
    errors of the network output in a plain rel-L2."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    mu = b.mean()
    return ((a - b).norm() / (b - mu).norm()).item()


# ---------------------------------------------------------------- oracle pins
def test_surrogate_oracle_matches_reference_script(golden):
    g = golden("surrogate.pt")
    traj = surrogate_traj(g["seed"], g["n"])
    assert abs(float(np.abs(traj).sum()) - g["traj_checksum"]) < 1e-3 * g["traj_checksum"]
    norm = O.Normalizer("gaussian", **g["norm"])
    fn = lambda x: O.fno3d_forward(g["sd"], x, g["ctor"][6])
    got = WO.materialize_surrogate(fn, norm, traj, g["gas_ratio"], g["equivalence_ratio"], g["step"], g["batch_size"],
                                   g["sub_s"])
    assert got.shape == tuple(g["pred_traj"].shape) == (9, 128, 128)
    assert centred_rel_l2(got, g["pred_traj"]) < TOL


@pytest.mark.parametrize("case", ["ft3d", "ft3d_clipped", "ft2d", "ft2d_clipped", "galerkin3d"])
def test_sibling_oracles_match_reference_modules(golden, case):
    g = golden("siblings.pt")[case]
    sd = g["sd"]
    if g["kind"] == "ft3d":
        y = WO.mwt_sparse_kernel_ft3d(g["x"], [sd[f"weights{k}"] for k in (1, 2, 3, 4)], g["alpha"], sd["Lo.weight"],
                                      sd["Lo.bias"])
    elif g["kind"] == "ft2d":
        y = WO.mwt_sparse_kernel_ft2d(g["x"], [sd[f"weights{k}"] for k in (1, 2)], g["alpha"], sd["Lo.weight"],
                                      sd["Lo.bias"])
    else:
        y = WO.galerkin_spectral_conv3d(g["x"], [sd[f"weights{k}"] for k in (1, 2, 3, 4)])
    assert y.shape == g["y"].shape
    assert O.rel_l2(y, g["y"]) < TOL


# ---------------------------------------------------------------- surrogate host logic
def test_window_plan_follows_the_script_loop():
    from realpdebench_b200.surrogate import window_plan
    # 9 frames, step 2, 2 windows per forward: range(0, 8, 4) -> [0:4], [4:8]; tail [7:9]; 8 + 1 predictions
    assert window_plan(9, 2, 2) == ([(0, 4), (4, 8)], (7, 9), 9)
    # the real script: 10-frame windows, 50 per forward, 1001 frames -> two full chunks + the tail frame
    chunks, tail, n_pred = window_plan(1001, 10, 50)
    assert chunks == [(0, 500), (500, 1000)] and tail == (991, 1001) and n_pred == 1001
    # a short last chunk that is still a whole number of windows is fine (reshape(-1, step, ...))
    assert window_plan(30, 10, 2) == ([(0, 20), (20, 30)], (20, 30), 31)
    # ... and one that is not fails like the reference's reshape (generate_surrogate_data.py:65)
    with pytest.raises(RuntimeError, match="invalid for a chunk of 5 frames"):
        window_plan(25, 10, 2)
    with pytest.raises(RuntimeError, match="fewer than one window"):
        window_plan(3, 10, 2)
    # oracle and plan agree on the number of predicted frames
    for n, step, bs in ((9, 2, 2), (13, 3, 2), (21, 10, 1)):
        traj = np.zeros((n, 4, 4, 2), np.float32)
        out = WO.materialize_surrogate(lambda x: x[..., :1], O.Normalizer("none"), traj, 1, 1, step, bs)
        assert out.shape[0] == window_plan(n, step, bs)[2]


class _HostPipes:
    """Stand-in for surrogate._CudaPipes: same call protocol, host memory, records the call order."""

    def __init__(self):
        self.log = []
        self.events = 0

    def pinned(self, numel):
        return torch.empty(numel, dtype=torch.float32)

    def h2d(self, host):
        self.events += 1
        self.log.append(("h2d", host.numel()))
        return host.clone(), self.events

    def host_wait(self, ev):
        self.log.append(("host_wait", ev))

    def compute_wait(self, ev, d):
        self.log.append(("compute_wait", ev))

    def d2h(self, host, p):
        self.log.append(("d2h", p.numel()))
        host.copy_(p.reshape(-1))

    def finish(self):
        self.log.append(("finish",))


class _OracleBackedModel(nn.Module):
    """What materialize_surrogate needs from the engine model: shape_in / shape_out / parameters / rollout(x,a,b,1)."""

    def __init__(self, sd, shape_in, shape_out):
        super().__init__()
        self.sd, self.shape_in, self.shape_out = sd, tuple(shape_in), tuple(shape_out)
        self.p = nn.Parameter(torch.zeros(1))
        self.batches = []

    def rollout(self, x, a, b, n_steps):
        assert n_steps == 1 and not torch.is_grad_enabled()
        self.batches.append(x.shape[0])
        return O.fno3d_forward(self.sd, x, self.shape_out) * a + b


@pytest.mark.parametrize("kind", ["gaussian", "range", "none"])
def test_materialize_surrogate_host_logic_matches_oracle(kind):
    from realpdebench_b200.surrogate import materialize_surrogate
    torch.manual_seed(90)
    step, h0, w0, c = 3, 12, 10, 4
    s_in, s_out = (step, 6, 5, c + 2), (step, 6, 5, 1)  # sub_s = 2
    sd = O.init_state(3, (2, 2, 2), 2, 6, s_in, s_out)
    O.randomize_bn(sd, 91)
    norm = O.synthetic_normalizer(c + 2, 1, seed=92, kind=kind)
    traj = torch.randn(13, h0, w0, c).double().numpy()  # float64 on disk -> float32 like torch.tensor(.., dtype=float)
    want = WO.materialize_surrogate(lambda x: O.fno3d_forward(sd, x, s_out), norm, traj, 60, 1.1, step, 2, 2)
    model, pipes = _OracleBackedModel(sd, s_in, s_out), _HostPipes()
    got = materialize_surrogate(model, norm, traj, 60, 1.1, step=step, batch_size=2, sub_s=2, _pipes=pipes)
    assert got.shape == want.shape == (13, 6, 5) and got.dtype == np.float32
    assert O.rel_l2(torch.from_numpy(got), torch.from_numpy(want)) < TOL
    assert model.batches == [2, 2, 1]  # two chunks of two windows, then the tail window
    kinds = [e[0] for e in pipes.log]
    # the next chunk is staged only after the current forward was enqueued; buffer reuse waits on its last copy
    assert kinds == ["h2d", "compute_wait", "d2h", "h2d", "compute_wait", "d2h", "host_wait", "h2d", "compute_wait",
                     "d2h", "finish"]
    assert pipes.log[6] == ("host_wait", 1)  # window 2 reuses the staging buffer of window 0


def test_materialize_surrogate_rejects_cpu_models_and_wrong_shapes():
    from realpdebench_b200.surrogate import materialize_surrogate
    sd = O.init_state(3, (2, 2, 2), 1, 4, (2, 4, 4, 3), (2, 4, 4, 1))
    model = _OracleBackedModel(sd, (2, 4, 4, 3), (2, 4, 4, 1))
    traj = np.zeros((5, 4, 4, 1), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        materialize_surrogate(model, O.Normalizer("none"), traj, 1, 1, step=2, batch_size=1)
    with pytest.raises(RuntimeError, match="does not match windows"):
        materialize_surrogate(model, O.Normalizer("none"), np.zeros((5, 4, 4, 2), np.float32), 1, 1, step=2,
                              batch_size=1, _pipes=_HostPipes())
    with pytest.raises(ValueError, match=r"\[n, H, W, C\]"):
        materialize_surrogate(model, O.Normalizer("none"), traj[0], 1, 1, step=2, batch_size=1)


def test_postprocess_affine_is_the_target_denormalisation():
    from realpdebench_b200.surrogate import postprocess_affine
    p = torch.randn(2, 3, 4, 5, 2)
    for kind in ("gaussian", "range", "none"):
        norm = O.synthetic_normalizer(6, 2, seed=7, kind=kind)
        a, b = postprocess_affine(norm, 2, "cpu")
        assert torch.allclose(p * a + b, norm.postprocess(p, p)[1], atol=1e-6)


# ---------------------------------------------------------------- sibling routing
class sparseKernelFT2d(nn.Module):
    """Stand-in with the reference class name / attributes (the GPU box has no reference package)."""

    def __init__(self, sd, modes):
        super().__init__()
        self.modes = modes
        self.weights1, self.weights2 = nn.Parameter(sd["weights1"].clone()), nn.Parameter(sd["weights2"].clone())
        self.Lo = nn.Linear(sd["Lo.weight"].shape[1], sd["Lo.weight"].shape[0])
        with torch.no_grad():
            self.Lo.weight.copy_(sd["Lo.weight"]), self.Lo.bias.copy_(sd["Lo.bias"])
        self.reference_calls = 0

    def forward(self, x):
        self.reference_calls += 1
        return WO.mwt_sparse_kernel_ft2d(x, [self.weights1, self.weights2], self.modes, self.Lo.weight, self.Lo.bias)


def test_route_keeps_reference_forward_for_cpu_and_autograd_calls(golden):
    from realpdebench_b200 import siblings
    g = golden("siblings.pt")["ft2d"]
    m = sparseKernelFT2d(g["sd"], g["alpha"])
    assert siblings.engine_forward_for(m) is siblings.mwt_sparse_kernel_ft2d_forward
    assert siblings.route(m) and siblings.route(m)  # idempotent
    with torch.no_grad():
        y = m(g["x"])  # CPU tensor: the module's own forward, untouched
    assert m.reference_calls == 1 and O.rel_l2(y, g["y"]) < TOL
    y = m(g["x"])  # autograd on: reference forward (the engine operator has no backward)
    assert m.reference_calls == 2 and y.requires_grad
    assert siblings.route_all(nn.Sequential(m, nn.ReLU())) == 1
    assert not siblings.route(nn.ReLU())
    siblings.unroute(m)
    assert "forward" not in m.__dict__
    with torch.no_grad():
        m(g["x"])
    assert m.reference_calls == 3


def test_engine_sibling_forward_fails_loudly_on_cpu(golden):
    from realpdebench_b200 import siblings
    g = golden("siblings.pt")["ft2d"]
    m = sparseKernelFT2d(g["sd"], g["alpha"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        siblings.mwt_sparse_kernel_ft2d_forward(m, g["x"])
