"""B200-native FNO forward / rollout engine behind RealPDEBench's model registry.

Public surface (mirrors the reference names):

* ``FNO3d`` / ``SpectralConv3d``  - realpdebench/model/fno.py
* ``FNO2d`` / ``SpectralConv2d``  - the 2-D variant defined in SURVEY.md 8(c)
* ``load_model``                  - realpdebench/model/load_model.py (+ ``fno2d``)
* ``rollout``                     - realpdebench/eval.py:296-326 for one batch
* ``rollout_stream``              - the same over a DataLoader, H2D of batch i+1 overlapped with batch i
* ``install``                     - registers the above at ``realpdebench.model.fno``
                                    so the unmodified reference scripts use them
* ``eval_metrics``                - realpdebench/utils/metrics.py:24-131 on the GPU
* ``python -m realpdebench_b200.run {train,eval,train_surrogate} ...`` - ``install()`` + the unmodified reference script
* ``materialize_surrogate``       - realpdebench/data/generate_surrogate_data.py:58-88 for one trajectory
* ``siblings``                    - the spectral operator behind the MWT / Galerkin ``bixyz,ioxyz->boxyz`` layers
* ``dist`` / ``optim``            - gradient all-reduce under the backward pass, fused Adam (training path)

The arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of
``include/b200fno.h``; there is no CPU or PyTorch fallback.
"""
from .fno import FNO2d, FNO3d, SpectralConv2d, SpectralConv3d
from .install import install, uninstall
from .load_model import load_model
from .metrics import eval_metrics
from .rollout import rollout, rollout_affine, rollout_stream
from .surrogate import materialize_surrogate

__all__ = ["FNO3d", "FNO2d", "SpectralConv3d", "SpectralConv2d", "load_model", "rollout", "rollout_affine", "rollout_stream",
           "install", "uninstall", "eval_metrics", "materialize_surrogate"]
__version__ = "0.1.0"
