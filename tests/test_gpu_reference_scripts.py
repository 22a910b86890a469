"""SURVEY 8f row N2 on the GPU: the UNMODIFIED reference ``train.py`` and ``eval.py`` complete a run on the engine.

The reference package travels to the GPU box as ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref`` of the
reference checkout); nothing here reads ``/root/reference``.  On a synthetic HF-Arrow cylinder data
set in the reference's own on-disk format (``data/fluid_hf_dataset.py:130-180``) and its ``configs/cylinder/fno.yaml``
(only paths, worker / batch counts, ``num_update`` and the probe diagnostic changed in a temporary copy):

1. ``python -m realpdebench_b200.run train ...`` trains 100 iterations on ``cuda:0`` through ``install()`` - the script's own
   loop (train.py:321-334) drives the engine's training forward / backward, validates with the CUDA ``eval_metrics`` and
   writes checkpoints in the reference format;
2. ``python -m realpdebench_b200.run eval ...`` loads the last checkpoint and runs the 10-step rollout loop
   (eval.py:296-352) on the engine;
3. the reference ALONE (no ``install()``, ``CUDA_VISIBLE_DEVICES=""``, its own PyTorch ``FNO3d`` on the host cores)
   evaluates the same checkpoint: every metric of its "Test results" line must agree with the engine's run to the
   printed precision (5 decimals) or 1e-4 relative.
"""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFPKG = os.path.join(ROOT, "baseline", "_ref")
STUBS = "matplotlib.pyplot,h5py"

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REFPKG, "realpdebench")),
                               reason="baseline/_ref (installed reference package) not present")


def _write_split(root, dtype, sims, t, h, w, splits):
    from datasets import Dataset
    rng = np.random.default_rng(0)
    rows = {"sim_id": [], "u": [], "v": [], "p": [], "shape_t": [], "shape_h": [], "shape_w": []}
    for s in sims:
        # smooth travelling waves + noise: something a network can fit a little, so that the metrics are not all ~1
        tt, yy, xx = np.meshgrid(np.arange(t), np.arange(h), np.arange(w), indexing="ij")
        for i, k in enumerate(("u", "v", "p")):
            f = np.sin(0.3 * xx - 0.2 * tt + i) * np.cos(0.25 * yy + 0.1 * tt) + 0.05 * rng.standard_normal((t, h, w))
            rows[k].append(f.astype(np.float32).tobytes())
        rows["sim_id"].append(s), rows["shape_t"].append(t), rows["shape_h"].append(h), rows["shape_w"].append(w)
    hf = os.path.join(root, "cylinder", "hf_dataset")
    os.makedirs(hf, exist_ok=True)
    Dataset.from_dict(rows).save_to_disk(os.path.join(hf, dtype))
    for split in splits:
        idx = [{"sim_id": s, "time_id": tid} for s in sims for tid in (0, 5, 10)]
        with open(os.path.join(hf, f"{split}_index_{dtype}.json"), "w") as f:
            json.dump(idx, f)


def _env(cuda=True):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REFPKG, os.environ.get("PYTHONPATH", "")]))
    if not cuda:
        env["CUDA_VISIBLE_DEVICES"] = ""
    return env


def _engine(tmp_path, script, *args):
    cmd = [sys.executable, "-m", "realpdebench_b200.run", "--stub", STUBS, script, *args]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), timeout=900, env=_env())


def _reference_cpu(tmp_path, script, *args):
    """The reference script by itself on the host cores: same stubs for its optional imports, no install()."""
    boot = ("import sys, types, runpy\n"
            f"for name in {STUBS.split(',')!r}:\n"
            "    parts = name.split('.')\n"
            "    for i in range(1, len(parts) + 1):\n"
            "        mod = '.'.join(parts[:i])\n"
            "        sys.modules.setdefault(mod, types.ModuleType(mod))\n"
            "        if i > 1: setattr(sys.modules['.'.join(parts[:i - 1])], parts[i - 1], sys.modules[mod])\n"
            f"sys.argv = ['{script}.py'] + sys.argv[1:]\n"
            f"runpy.run_module('realpdebench.{script}', run_name='__main__')\n")
    return subprocess.run([sys.executable, "-c", boot, *args], capture_output=True, text=True, cwd=str(tmp_path),
                          timeout=1800, env=_env(cuda=False))


def _logs(path):
    out = []
    for dp, _, fs in os.walk(path):
        out += [os.path.join(dp, f) for f in fs if f.endswith(".log")]
    return "".join(open(p).read() for p in sorted(out, key=os.path.getmtime))


def _test_results(text):
    """metric name -> value from the LAST 'Test results:' record of eval.py:354-360."""
    tail = text[text.rindex("Test results:"):]
    tail = tail[:tail.index("Testing complete")] if "Testing complete" in tail else tail
    return {k.strip(): float(v) for k, v in re.findall(r"([a-z][a-z0-9 ]*?): (-?(?:\d+\.\d+|inf|nan))", tail)}


@needs_ref
def test_unmodified_train_and_eval_complete_on_the_engine_and_match_the_reference(tmp_path):
    import yaml
    root = str(tmp_path / "data")
    _write_split(root, "numerical", ["101.h5", "102.h5"], 60, 32, 48, ("train",))
    _write_split(root, "real", ["201.h5"], 240, 16, 24, ("train", "val", "test"))  # horizon = 20 + 10 * 20 frames
    with open(os.path.join(REFPKG, "realpdebench", "configs", "cylinder", "fno.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg.update(dataset_root=root, num_workers=0, results_path=str(tmp_path / "results"), train_batch_size=4,
               test_batch_size=2, num_update=100, is_use_tb=False, probe_diagnostic=False, N_plot=0, lr=1e-3)
    cfg.pop("checkpoint_path", None)
    cfg_path = str(tmp_path / "fno.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)

    # 1. train.py on the engine
    r = _engine(tmp_path, "train", "--config", cfg_path, "--use_hf_dataset")
    assert r.returncode == 0, r.stderr[-4000:]
    text = _logs(str(tmp_path / "results"))
    assert "Start training on cuda:0" in text
    assert "Number of parameters: 50357955" in text
    ckpts = sorted((os.path.join(dp, f) for dp, _, fs in os.walk(str(tmp_path / "results")) for f in fs
                    if f.endswith(".pth")), key=os.path.getmtime)
    assert ckpts, "train.py wrote no checkpoint"
    ck = torch.load(ckpts[-1], map_location="cpu", weights_only=False)
    losses = ck["train_losses"]
    assert len(losses) >= 2 and all(np.isfinite(losses)) and losses[-1] < losses[0], losses[:3] + losses[-3:]
    for k, v in ck["model_state_dict"].items():
        assert torch.isfinite(torch.view_as_real(v) if v.is_complex() else v.float()).all(), k
    # validation metrics stored by train.py must be plain host values (a checkpoint must load on a CPU-only box)
    for k, v in ck["val_losses"].items():
        assert all((not torch.is_tensor(x)) or x.device.type == "cpu" for x in v), k

    # 2. eval.py on the engine, 3. the reference alone on the CPU, same checkpoint
    res_e = str(tmp_path / "results_engine")
    res_r = str(tmp_path / "results_ref")
    for res in (res_e, res_r):
        c2 = dict(cfg, results_path=res)
        with open(res + ".yaml", "w") as f:
            yaml.safe_dump(c2, f)
    r = _engine(tmp_path, "eval", "--config", res_e + ".yaml", "--use_hf_dataset", "--checkpoint_path", ckpts[-1])
    assert r.returncode == 0, r.stderr[-4000:]
    te = _logs(res_e)
    assert "Start testing on cuda:0" in te and "Testing complete" in te
    r = _reference_cpu(tmp_path, "eval", "--config", res_r + ".yaml", "--use_hf_dataset", "--checkpoint_path", ckpts[-1])
    assert r.returncode == 0, r.stderr[-4000:]
    tr = _logs(res_r)
    assert "Start testing on cpu" in tr and "Testing complete" in tr
    me, mr = _test_results(te), _test_results(tr)
    assert set(me) == set(mr) and len(mr) >= 14, (sorted(me), sorted(mr))
    for k, v in mr.items():
        if np.isfinite(v):
            assert abs(me[k] - v) <= 2e-5 + 1e-4 * abs(v), (k, me[k], v)
        else:
            assert str(me[k]) == str(v), (k, me[k], v)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "n2_eval_metrics.json"), "w") as f:
        json.dump({"engine_cuda": me, "reference_cpu": mr, "train_losses_first_last": [losses[0], losses[-1]]}, f)
