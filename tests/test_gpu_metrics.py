"""GPU parity of eval_metrics (realpdebench/utils/metrics.py:24-131) through the C-ABI (b200fno_eval_metrics): against
the golden outputs recorded from the reference function itself (tests/golden/metrics.pt) and against the vectorised
oracle on larger random fields.  Tolerance: 2e-4 relative per scalar (fp32 reductions over up to ~1e6 elements and fp32
DFT sums of up to a few hundred terms, in a different order than torch's)."""
import pytest
import torch

from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def close(got, ref, tol=2e-4):
    got, ref = torch.stack([g.cpu() for g in got]).double(), torch.as_tensor(ref).double()
    for i, name in enumerate(M.NAMES):
        a, b = float(got[i]), float(ref[i])
        if b != b or abs(b) == float("inf"):
            assert (a != a) if b != b else a == b, (name, a, b)
        else:
            assert abs(a - b) <= tol * max(abs(b), 1e-6), (name, a, b)


def test_metrics_match_reference_golden(golden):
    from realpdebench_b200.metrics import eval_metrics
    for case in golden("metrics.pt"):
        got = eval_metrics(case["pred"].to(dev()), case["target"].to(dev()), case["c"], case["batch_size"])
        close(got, case["out"])


@pytest.mark.parametrize("shape,c,bs", [
    ((4, 24, 20, 28, 3), 3, None),      # nb = 10, three channels
    ((6, 40, 64, 128, 3), 2, 4),        # cylinder-like grid, real-data convention (2 of 3 channels), 2 chunks
    ((2, 9, 7, 11, 5), 5, None),        # odd sizes
])
def test_metrics_match_oracle(shape, c, bs):
    from realpdebench_b200.metrics import eval_metrics
    torch.manual_seed(70)
    target = torch.randn(*shape) + torch.linspace(0, 2, shape[1]).reshape(1, -1, 1, 1, 1)
    pred = target + 0.2 * torch.randn(*shape)
    ref = torch.stack(M.eval_metrics(pred, target, c, bs))
    close(eval_metrics(pred.to(dev()), target.to(dev()), c, bs), ref)


def test_metrics_errors():
    from realpdebench_b200.metrics import eval_metrics
    x = torch.randn(2, 4, 4, 4, 2)
    with pytest.raises(RuntimeError):
        eval_metrics(x.to(dev()), x.to(dev()), 3)  # more channels than present
    with pytest.raises(RuntimeError):
        eval_metrics(x.to(dev()), x[:1].to(dev()), 2)  # shape mismatch


def test_metrics_accept_host_tensors(golden):
    """eval.py:342-352 hands CPU tensors (pred.cpu()) to eval_metrics: chunks are staged to the GPU one at a time."""
    from realpdebench_b200.metrics import eval_metrics
    case = golden("metrics.pt")[1]
    close(eval_metrics(case["pred"], case["target"], case["c"], case["batch_size"]), case["out"])
