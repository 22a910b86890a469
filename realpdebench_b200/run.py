"""Launcher: run an UNMODIFIED reference script on the engine.

    python -m realpdebench_b200.run train --config configs/cylinder/fno.yaml --use_hf_dataset
    python -m realpdebench_b200.run eval  --config configs/cylinder/fno.yaml --checkpoint_path model.pth
    python -m realpdebench_b200.run train_surrogate --config configs/combustion/surrogate_model/fno.yaml

``install()`` (the engine's module at ``sys.modules['realpdebench.model.fno']``, ``fno2d`` in the registry,
GPU ``eval_metrics``) and then ``runpy`` of ``realpdebench.<script>`` as ``__main__`` with the remaining arguments - the
script files, ``load_model.py`` and the YAML configs are used as they are.

``--stub NAME[,NAME...]`` registers empty modules for optional imports of the reference that a minimal environment
lacks (``realpdebench/utils/metrics.py:8`` imports ``matplotlib.pyplot``, the HDF5 datasets import ``h5py``); they are
only ever called by the plotting / HDF5 code paths, not by the FNO path.
"""
from __future__ import annotations

import runpy
import sys
import types

SCRIPTS = ("train", "eval", "train_surrogate")


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    stubs = []
    while argv and argv[0] == "--stub":
        if len(argv) < 2:
            raise SystemExit("--stub needs a comma-separated module list")
        stubs += [s for s in argv[1].split(",") if s]
        argv = argv[2:]
    if not argv or argv[0] not in SCRIPTS:
        raise SystemExit(f"usage: python -m realpdebench_b200.run [--stub MOD,...] {{{','.join(SCRIPTS)}}} [script args]")
    script, rest = argv[0], argv[1:]
    for name in stubs:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            mod = ".".join(parts[:i])
            if mod not in sys.modules:
                sys.modules[mod] = types.ModuleType(mod)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[mod])
    try:
        import realpdebench  # noqa: F401
    except ImportError as e:
        raise SystemExit(f"realpdebench_b200.run: the reference package 'realpdebench' is not importable ({e}); "
                         "install it or put its checkout on PYTHONPATH")
    from . import install
    install()
    sys.argv = [f"{script}.py"] + rest
    runpy.run_module(f"realpdebench.{script}", run_name="__main__")


if __name__ == "__main__":
    main()
