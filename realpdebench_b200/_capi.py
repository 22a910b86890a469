"""ctypes binding of ``include/b200fno.h`` (the C-ABI of the CUDA engine).

There is no fallback: if the shared library has not been built, importing the
symbols raises, and every compute entry point needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

ABI_VERSION = 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
COMPUTE = {"f32": 0, "bf16": 1}  # B200FNO_COMPUTE_*

# every extern "C" symbol declared in include/b200fno.h
SYMBOLS = (
    "b200fno_last_error", "b200fno_abi_version", "b200fno_plan_create", "b200fno_plan_destroy",
    "b200fno_plan_set_impl", "b200fno_plan_get_impl", "b200fno_plan_set_compute", "b200fno_plan_get_compute", "b200fno_plan_workspace_bytes", "b200fno_plan_packed_bytes",
    "b200fno_plan_bind", "b200fno_pack_weights", "b200fno_forward", "b200fno_rollout",
    "b200fno_spectral_workspace_bytes", "b200fno_spectral_conv", "b200fno_launch_count",
    "b200fno_debug_first_nonfinite", "b200fno_spectral_cache_clear",
    "b200fno_launch_count_reset", "b200fno_host_table", "b200fno_host_table_slice", "b200fno_algorithmic_bytes", "b200fno_timing_enable",
    "b200fno_timing_collect", "b200fno_selftest_umma", "b200fno_selftest_mma_rate",
    "b200fno_train_workspace_bytes", "b200fno_train_bind", "b200fno_train_forward", "b200fno_train_backward",
    "b200fno_adam_step", "b200fno_metrics_workspace_bytes", "b200fno_eval_metrics", "b200fno_plan_stage_impl",
)


class B200FNOError(RuntimeError):
    """Non-zero status from the C-ABI (mirrors TORCH_CHECK -> RuntimeError in the reference's own extension)."""


class Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "ndim", "max_batch", "t_in", "t_out", "h", "w", "c_in", "c_out", "width", "n_layers",
        "modes1", "modes2", "modes3", "padding", "proj_hidden")] + [("bn_eps", C.c_float)]


_fp = C.c_void_p  # device pointers travel as integers
_fpp = C.POINTER(C.c_void_p)


class Weights(C.Structure):
    _fields_ = [("fc0_w", _fp), ("fc0_b", _fp), ("spec_w", _fpp), ("conv_w", _fpp), ("conv_b", _fpp),
                ("bn_weight", _fpp), ("bn_bias", _fpp), ("bn_mean", _fpp), ("bn_var", _fpp),
                ("fc1_w", _fp), ("fc1_b", _fp), ("fc2_w", _fp), ("fc2_b", _fp)]


class Grads(C.Structure):
    _fields_ = [("fc0_w", _fp), ("fc0_b", _fp), ("spec_w", _fpp), ("conv_w", _fpp), ("conv_b", _fpp),
                ("bn_weight", _fpp), ("bn_bias", _fpp), ("fc1_w", _fp), ("fc1_b", _fp), ("fc2_w", _fp), ("fc2_b", _fp),
                ("x", _fp)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the b200fno CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    i32, i64, sz, vp = C.c_int32, C.c_int64, C.c_size_t, C.c_void_p
    L.b200fno_last_error.restype = C.c_char_p
    L.b200fno_last_error.argtypes = []
    L.b200fno_abi_version.restype = C.c_int
    L.b200fno_plan_create.restype = C.c_int
    L.b200fno_plan_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    L.b200fno_plan_destroy.restype = C.c_int
    L.b200fno_plan_destroy.argtypes = [vp]
    L.b200fno_plan_set_impl.restype = C.c_int
    L.b200fno_plan_set_impl.argtypes = [vp, C.c_int]
    L.b200fno_plan_set_compute.restype = C.c_int
    L.b200fno_plan_set_compute.argtypes = [vp, C.c_int]
    L.b200fno_plan_get_compute.restype = C.c_int
    L.b200fno_plan_get_compute.argtypes = [vp]
    L.b200fno_plan_get_impl.restype = C.c_int
    L.b200fno_plan_get_impl.argtypes = [vp]
    L.b200fno_plan_workspace_bytes.restype = sz
    L.b200fno_plan_workspace_bytes.argtypes = [vp]
    L.b200fno_plan_packed_bytes.restype = sz
    L.b200fno_plan_packed_bytes.argtypes = [vp]
    L.b200fno_plan_bind.restype = C.c_int
    L.b200fno_plan_bind.argtypes = [vp, vp, sz, vp, sz]
    L.b200fno_pack_weights.restype = C.c_int
    L.b200fno_pack_weights.argtypes = [vp, C.POINTER(Weights), vp]
    L.b200fno_forward.restype = C.c_int
    L.b200fno_forward.argtypes = [vp, i32, vp, vp, vp]
    L.b200fno_rollout.restype = C.c_int
    L.b200fno_rollout.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp]
    L.b200fno_spectral_workspace_bytes.restype = sz
    L.b200fno_spectral_workspace_bytes.argtypes = [i32] * 10
    L.b200fno_spectral_conv.restype = C.c_int
    L.b200fno_spectral_conv.argtypes = [i32] * 10 + [_fpp, vp, vp, vp, sz, vp]
    L.b200fno_spectral_cache_clear.restype = None
    L.b200fno_spectral_cache_clear.argtypes = []
    L.b200fno_debug_first_nonfinite.restype = C.c_int
    L.b200fno_debug_first_nonfinite.argtypes = [C.c_void_p]
    L.b200fno_launch_count.restype = i64
    L.b200fno_launch_count.argtypes = []
    L.b200fno_launch_count_reset.restype = None
    L.b200fno_launch_count_reset.argtypes = []
    L.b200fno_host_table.restype = i64
    L.b200fno_host_table.argtypes = [i32] * 8 + [C.POINTER(C.c_float), i64, C.POINTER(i32), C.POINTER(i32),
                                                 C.POINTER(i32)]
    L.b200fno_host_table_slice.restype = i64
    L.b200fno_host_table_slice.argtypes = [i32] * 9 + [C.POINTER(C.c_float), i64, C.POINTER(i32), C.POINTER(i32),
                                                       C.POINTER(i32)]
    L.b200fno_timing_enable.restype = C.c_int
    L.b200fno_timing_enable.argtypes = [vp, C.c_int]
    L.b200fno_timing_collect.restype = C.c_int
    L.b200fno_timing_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.b200fno_selftest_umma.restype = C.c_int
    L.b200fno_selftest_umma.argtypes = [i32] * 5 + [vp, vp, vp, vp]
    L.b200fno_selftest_mma_rate.restype = C.c_int
    L.b200fno_selftest_mma_rate.argtypes = [i32, i32, i32, i32, i32, vp, vp]
    L.b200fno_train_workspace_bytes.restype = sz
    L.b200fno_train_workspace_bytes.argtypes = [vp]
    L.b200fno_train_bind.restype = C.c_int
    L.b200fno_train_bind.argtypes = [vp, vp, sz]
    L.b200fno_train_forward.restype = C.c_int
    L.b200fno_train_forward.argtypes = [vp, i32, vp, vp, _fpp, _fpp, C.c_float, vp]
    L.b200fno_train_backward.restype = C.c_int
    L.b200fno_train_backward.argtypes = [vp, i32, vp, vp, C.POINTER(Grads), _fpp, vp]
    L.b200fno_adam_step.restype = C.c_int
    L.b200fno_adam_step.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp]
    L.b200fno_metrics_workspace_bytes.restype = sz
    L.b200fno_metrics_workspace_bytes.argtypes = [i32] * 6
    L.b200fno_eval_metrics.restype = C.c_int
    L.b200fno_eval_metrics.argtypes = [vp, vp] + [i32] * 6 + [vp, sz, vp, vp]
    L.b200fno_plan_stage_impl.restype = C.c_int
    L.b200fno_plan_stage_impl.argtypes = [vp, i32]
    L.b200fno_algorithmic_bytes.restype = C.c_double
    L.b200fno_algorithmic_bytes.argtypes = [vp, i32]
    if L.b200fno_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.b200fno_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


STAGES = ("lift", "fwdW", "fwdH", "fwdT", "modes", "invT", "invH", "layer", "proj")


def check(rc: int) -> None:
    if rc != 0:
        raise B200FNOError(f"b200fno error {rc}: {lib().b200fno_last_error().decode(errors='replace')}")


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr


def host_table(ndim, t, h, w, m1, m2, m3, which, kw0=0):
    """numpy copy of a truncated-DFT table + (ld, kept T freqs, kept H freqs); ``kw0``: first W frequency of a mode
    slice of ``m3`` frequencies (b200fno_host_table_slice)."""
    import numpy as np
    L = lib()
    ld = C.c_int32(0)
    ft = (C.c_int32 * max(t, 1))()
    fh = (C.c_int32 * max(h, 1))()
    n = L.b200fno_host_table_slice(ndim, t, h, w, m1, m2, m3, kw0, which, None, 0, C.byref(ld), ft, fh)
    if n < 0:
        check(int(n))
    buf = np.zeros(int(n), dtype=np.float32)
    L.b200fno_host_table_slice(ndim, t, h, w, m1, m2, m3, kw0, which, buf.ctypes.data_as(C.POINTER(C.c_float)), n,
                               C.byref(ld), ft, fh)
    kt = min(2 * m1, t) if ndim == 3 else 1
    kh = min(2 * m2, h)
    return buf.reshape(-1, ld.value), list(ft[:kt]), list(fh[:kh])
