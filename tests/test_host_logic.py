"""CPU tests of the host side: C-ABI surface, module/state_dict contract, registry drop-in."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import fno_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200fno.h")


def test_library_exports_every_declared_symbol():
    from realpdebench_b200 import _capi
    lib = _capi.lib()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(b200fno_[a-z_0-9]+)\s*\(", text))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.b200fno_abi_version() == _capi.ABI_VERSION


def test_header_is_plain_c_and_struct_layout_matches_binding(tmp_path):
    from realpdebench_b200 import _capi
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "b200fno.h"\nint main(void){printf("%zu %zu %zu\\n",'
                   'sizeof(b200fno_desc_t),sizeof(b200fno_weights_t),offsetof(b200fno_desc_t,bn_eps));return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_capi.Desc)
    assert int(out[1]) == ctypes.sizeof(_capi.Weights)
    assert int(out[2]) == _capi.Desc.bn_eps.offset


def test_no_device_errors_are_reported_not_crashed():
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from realpdebench_b200 import _capi
    lib = _capi.lib()
    d = _capi.Desc(abi_version=1, ndim=3, max_batch=1, t_in=4, t_out=4, h=8, w=8, c_in=2, c_out=2, width=8,
                   n_layers=1, modes1=2, modes2=2, modes3=2, padding=6, proj_hidden=128, bn_eps=1e-5)
    plan = ctypes.c_void_p()
    rc = lib.b200fno_plan_create(ctypes.byref(d), ctypes.byref(plan))
    assert rc == -3 and b"CUDA" in lib.b200fno_last_error()
    d.ndim = 4
    assert lib.b200fno_plan_create(ctypes.byref(d), ctypes.byref(plan)) == -1


def test_module_matches_reference_state_dict_and_init(golden):
    import realpdebench_b200 as R
    g = golden("kat_a.pt")
    torch.manual_seed(0)
    m = R.FNO3d(*g["ctor"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["sd"].keys())
    for k, v in g["sd"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
        if ".bns." not in "." + k:
            assert torch.equal(sd[k], v), k  # same RNG order as fno.py:89-103
    assert sum(p.numel() for p in m.parameters()) == sum(
        v.numel() for k, v in g["sd"].items() if "running" not in k and "num_batches" not in k)
    m.load_state_dict(g["sd"])


def test_cpu_forward_fails_loudly_no_fallback(golden):
    import realpdebench_b200 as R
    g = golden("kat_a.pt")
    m = R.FNO3d(*g["ctor"]).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(g["x"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.SpectralConv3d(2, 2, 1, 1, 1)(torch.randn(1, 2, 4, 4, 4))


def test_fno2d_module_contract():
    import realpdebench_b200 as R
    torch.manual_seed(4)
    s = (3, 10, 12, 2)
    m = R.FNO2d(4, 5, 2, 8, s, s)
    torch.manual_seed(4)
    sd = O.init_state(2, (4, 5), 2, 8, s, s)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_rollout_affine_is_post_then_pre():
    from realpdebench_b200 import rollout_affine
    for kind in ("gaussian", "range", "none"):
        n = O.synthetic_normalizer(5, 3, kind=kind)
        a, b = rollout_affine(n, 5, 3, "cpu")
        p = torch.randn(2, 4, 3)
        x = torch.randn(2, 4, 5)
        want = n.preprocess(n.postprocess(x, p)[1], p)[0]
        assert torch.allclose(p * a + b, want, atol=1e-6)


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_install_serves_unmodified_reference_registry():
    code = r'''
import sys, types, torch
for n in ("matplotlib", "matplotlib.pyplot", "h5py"):
    sys.modules.setdefault(n, types.ModuleType(n))
sys.path.insert(0, %r); sys.path.insert(0, %r)
import realpdebench_b200 as R
R.install()
from realpdebench.model.load_model import load_model
import yaml, os
cfg = yaml.safe_load(open(os.path.join(%r, "realpdebench/configs/cylinder/fno.yaml")))
ds = [(torch.zeros(20, 16, 16, 3), torch.zeros(20, 16, 16, 3))]
torch.manual_seed(0)
m = load_model(ds, "cpu", **cfg)
assert type(m).__module__ == "realpdebench_b200.fno", type(m)
from realpdebench.model.model import Model
assert isinstance(m, Model)
cfg["model_name"] = "fno2d"
m2 = load_model(ds, "cpu", **cfg)
assert type(m2).__name__ == "FNO2d" and m2.modes1 == cfg["modes2"] and m2.modes2 == cfg["modes3"]
R.uninstall()
from realpdebench.model.fno import FNO3d as RefFNO
assert RefFNO.__module__ == "realpdebench.model.fno"
torch.manual_seed(0)
ref = RefFNO(4, 12, 16, 4, 64, (20, 16, 16, 3), (20, 16, 16, 3))
a, b = m.state_dict(), ref.state_dict()
assert list(a) == list(b)
assert all(torch.equal(a[k], b[k]) for k in a)
try:
    load_model(ds, "cpu", model_name="nope")
except ValueError as e:
    assert "not supported" in str(e)
print("OK")
''' % (REF, ROOT, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


# ---------------------------------------------------------------- training / metrics host logic (no device needed)
def test_fused_adam_and_metrics_fail_loudly_on_cpu():
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from realpdebench_b200.metrics import eval_metrics
    from realpdebench_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.randn(4, 3))
    p.grad = torch.randn(4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdam([p], lr=1e-3).step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    x = torch.randn(2, 4, 4, 4, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eval_metrics(x, x, 2)
    with pytest.raises(RuntimeError, match="expects two"):
        eval_metrics(x, x[:1], 2)


def test_gradient_groups_follow_the_backward_order():
    """dist.OverlappedGradientReducer groups parameters the way b200fno_train_backward finishes them: projection
    first, then the Fourier layers from last to first, fc0 last - every parameter in exactly one group."""
    import realpdebench_b200 as R
    from realpdebench_b200 import dist as D
    m = R.FNO3d(2, 3, 3, 3, 8, (3, 7, 9, 2), (3, 7, 9, 2))
    names = [k for k, _ in m.named_parameters()]
    red = D.OverlappedGradientReducer.__new__(D.OverlappedGradientReducer)
    red.n_layers = m.n_layers
    groups = [red._group(names, idx) for idx in [m.n_layers] + list(range(m.n_layers - 1, -1, -1))]
    groups.append([n for n in names if n.startswith("fc0.")])
    assert groups[0] == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    assert all(".2." in n for n in groups[1]) and len(groups[1]) == 4 + 2 + 2  # 4 corners, conv w/b, bn w/b
    flat = [n for g in groups for n in g]
    assert sorted(flat) == sorted(names) and len(flat) == len(set(flat))


def test_install_routes_eval_metrics_only_with_cuda():
    code = (
        "import sys, types\n"
        "for n in ('matplotlib','matplotlib.pyplot','h5py'): sys.modules.setdefault(n, types.ModuleType(n))\n"
        "sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']\n"
        "sys.path.insert(0, '/root/reference')\n"
        "import torch, realpdebench_b200 as R\n"
        "import realpdebench.utils.metrics as M\n"
        "orig = M.eval_metrics\n"
        "R.install()\n"
        "patched = getattr(M.eval_metrics, '_b200fno_wrapped', False)\n"
        "assert patched == torch.cuda.is_available(), patched\n"
        "R.uninstall()\n"
        "assert M.eval_metrics is orig\n"
        "print('ok')\n")
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference package not present")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the engine arm): one JSON line with the
    contract's keys; under a 2-rank launch only rank 0 works and prints."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True, env=env, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fno_rollout_field_points_per_sec"
    assert d["unit"] == "field-points/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "fno2d_cylinder_256x512_rollout20" and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    # rank 1 of a multi-rank launch exits 0 without work or output
    r1 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, check=True, timeout=120,
                        env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.stdout.strip() == ""


def test_bench_engine_arm_weights_equal_cpu_arm_weights():
    """bench.py builds the engine model with the module's own init (no oracle import on the product path); the CPU
    arm uses oracle.init_state + randomize_bn.  Both must be the same tensors, bit for bit."""
    import realpdebench_b200 as R
    import bench
    for ndim, modes, s in ((2, (3, 4), (2, 12, 14, 3)), (3, (2, 3, 4), (4, 10, 12, 2))):
        m = bench.build_model(R, ndim, modes, 2, 8, s, s)
        sd = bench.build_state(ndim, modes, 2, 8, s, s)
        got = m.state_dict()
        assert set(got) == set(sd)
        for k, v in sd.items():
            assert torch.equal(got[k], v), k


def test_bench_loss_check_fixture_matches_bench_workload():
    import bench
    ref = bench.oracle_loss_check(bench.DEFAULT_WORKLOAD)
    ndim, modes, L, width, s_in, s_out, B, n_auto = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert ref is not None and ref["batch"] == B and ref["n_autoregressive"] == n_auto and ref["seed"] == 1234
    assert abs(ref["normalized_loss"] - sum(ref["per_sample"]) / B) < 1e-12


def test_compute_mode_selection_follows_autocast_and_rejects_unknown_modes():
    """Host logic of the bf16 mode: an active CUDA bf16 autocast context selects it, ``set_compute`` validates
    its argument and survives ``set_impl``; no device needed (the plan is only created on the first CUDA call)."""
    import realpdebench_b200 as R
    m = R.FNO2d(3, 4, 2, 16, (2, 8, 8, 2), (2, 8, 8, 2))
    assert m._resolve_compute() == ("f32", False) and m.engine.compute == "f32"
    if torch.cuda.is_available():  # torch disables a CUDA autocast context on a box without CUDA (covered by -m gpu)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            assert m._resolve_compute() == ("bf16", True)
        with torch.autocast("cuda", dtype=torch.float16):
            assert m._resolve_compute() == ("f32", False)  # only bf16 autocast has an engine mode
    m.set_compute("bf16")
    assert m._resolve_compute() == ("bf16", False) and m.engine.compute == "bf16"
    m.set_impl("simt")
    assert m.engine.compute == "bf16"  # a rebuilt engine keeps the mode
    with pytest.raises(ValueError, match="compute must be one of"):
        m.set_compute("fp8")
    from realpdebench_b200 import _capi
    lib = _capi.lib()
    assert lib.b200fno_plan_set_compute(None, 1) != 0 and "compute" in lib.b200fno_last_error().decode()
