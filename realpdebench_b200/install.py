"""Register the engine at the reference's registry boundary.

``realpdebench.model.load_model.load_model`` does a *lazy*
``from realpdebench.model.fno import FNO3d`` (load_model.py:12-13); placing this
package's ``fno`` module at ``sys.modules['realpdebench.model.fno']`` makes the
unmodified reference ``train.py`` / ``eval.py`` / ``train_surrogate.py`` build
the engine model from the unchanged ``configs/*/fno.yaml``.  ``load_model`` is
additionally wrapped so the new ``model_name: fno2d`` is served.
"""
import sys

_saved = {}


def install(wrap_load_model: bool = True, gpu_metrics: bool = True) -> None:
    import importlib
    engine_fno = importlib.import_module(__package__ + ".fno")
    engine_lm = importlib.import_module(__package__ + ".load_model")  # the submodule, not the re-exported function

    if "realpdebench.model.fno" not in _saved:
        _saved["realpdebench.model.fno"] = sys.modules.get("realpdebench.model.fno")
    sys.modules["realpdebench.model.fno"] = engine_fno
    try:
        import realpdebench.model as ref_model_pkg
        ref_model_pkg.fno = engine_fno
    except Exception:
        return  # reference package absent: engine classes are still importable from realpdebench_b200
    if wrap_load_model:
        try:
            import realpdebench.model.load_model as ref_lm
        except Exception:
            return
        if not getattr(ref_lm.load_model, "_b200fno_wrapped", False):
            _saved["load_model"] = ref_lm.load_model
            ref_lm.load_model = engine_lm.make_wrapper(ref_lm.load_model)
    if gpu_metrics:
        # eval.py:22 / train.py:21 do ``from realpdebench.utils.metrics import eval_metrics`` at import time: replacing
        # the attribute before they are imported routes the evaluation metrics (utils/metrics.py:24-131, two Python
        # triple loops on the CPU) to the CUDA kernels.  Without a CUDA device the reference function stays in place.
        try:
            import torch
            import realpdebench.utils.metrics as ref_metrics
        except Exception:
            return
        if torch.cuda.is_available() and not getattr(ref_metrics.eval_metrics, "_b200fno_wrapped", False):
            from .metrics import eval_metrics as gpu_eval_metrics

            def eval_metrics(pred, target, c, batch_size=None):
                # the reference returns its scalars on ``target.device`` (utils/metrics.py:36); train.py:376-416
                # torch.saves them into checkpoints, so host inputs must give host results
                out = gpu_eval_metrics(pred, target, c, batch_size)
                return tuple(v.to(target.device) for v in out)

            eval_metrics._b200fno_wrapped = True
            _saved["eval_metrics"] = ref_metrics.eval_metrics
            ref_metrics.eval_metrics = eval_metrics


def uninstall() -> None:
    if "realpdebench.model.fno" in _saved:
        prev = _saved.pop("realpdebench.model.fno")
        if prev is None:
            sys.modules.pop("realpdebench.model.fno", None)
        else:
            sys.modules["realpdebench.model.fno"] = prev
            try:
                import realpdebench.model as ref_model_pkg
                ref_model_pkg.fno = prev
            except Exception:
                pass
    if "load_model" in _saved:
        import realpdebench.model.load_model as ref_lm
        ref_lm.load_model = _saved.pop("load_model")
    if "eval_metrics" in _saved:
        import realpdebench.utils.metrics as ref_metrics
        ref_metrics.eval_metrics = _saved.pop("eval_metrics")
