"""Host-side owner of one C-ABI plan: buffers, packed weights, launches.

PyTorch is used for device memory (caching allocator) and the current stream;
every kernel is launched by ``libb200fno.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _capi
from ._capi import Desc, Grads, Weights, check


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"b200fno: {what} is on {t.device}; the engine runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"b200fno: {what} must be float32, got {t.dtype}")


class FNOEngine:
    """Plan + workspace + packed weights for one FNO module (eval-mode forward and rollout)."""

    def __init__(self, ndim: int, modes: Sequence[int], n_layers: int, width: int, shape_in: Sequence[int],
                 shape_out: Sequence[int], padding: int = 6, bn_eps: float = 1e-5, impl: str = "auto"):
        self.ndim = ndim
        self.modes = tuple(int(m) for m in modes)
        self.n_layers, self.width = int(n_layers), int(width)
        self.shape_in = tuple(int(s) for s in shape_in)
        self.shape_out = tuple(int(s) for s in shape_out)
        self.padding, self.bn_eps = int(padding), float(bn_eps)
        self.impl = impl
        self.compute = "f32"  # "f32" | "bf16" (torch.autocast(bfloat16) semantics), see set_compute
        self._plan: Optional[C.c_void_p] = None
        self._max_batch = 0
        self._device: Optional[torch.device] = None
        self._ws = self._packed = self._train_ws = None
        self._weights_key = None
        self._keepalive: List[torch.Tensor] = []
        self._train_seq = 0  # id of the train-mode forward whose activations the training workspace holds

    # -- plan management ------------------------------------------------------
    def _destroy(self):
        if self._plan is not None:
            _capi.lib().b200fno_plan_destroy(self._plan)
        self._plan, self._ws, self._weights_key, self._train_ws = None, None, None, None
        self.__dict__.pop("_graphs", None)  # captured rollouts hold the old plan's addresses

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure_plan(self, batch: int, device: torch.device):
        if self._plan is not None and batch <= self._max_batch and device == self._device:
            return
        self._destroy()
        L = _capi.lib()
        m = self.modes if self.ndim == 3 else (1, *self.modes[-2:])
        d = Desc(abi_version=_capi.ABI_VERSION, ndim=self.ndim, max_batch=batch, t_in=self.shape_in[0],
                 t_out=self.shape_out[0], h=self.shape_in[1], w=self.shape_in[2], c_in=self.shape_in[3],
                 c_out=self.shape_out[3], width=self.width, n_layers=self.n_layers, modes1=m[0], modes2=m[1],
                 modes3=m[2], padding=self.padding, proj_hidden=128, bn_eps=self.bn_eps)
        plan = C.c_void_p()
        with torch.cuda.device(device):
            check(L.b200fno_plan_create(C.byref(d), C.byref(plan)))
            self._plan = plan
            if self.impl != "auto":
                check(L.b200fno_plan_set_impl(plan, {"simt": _capi.IMPL_SIMT, "tc": _capi.IMPL_TC}[self.impl]))
            check(L.b200fno_plan_set_compute(plan, _capi.COMPUTE[self.compute]))
            wsb, pkb = L.b200fno_plan_workspace_bytes(plan), L.b200fno_plan_packed_bytes(plan)
            self._ws = torch.empty(wsb, dtype=torch.uint8, device=device)
            if os.environ.get("B200FNO_POISON_WS"):  # debugging aid: every fp32 word of the workspace starts as NaN
                self._ws.fill_(0xFF)
            if self._packed is None or self._packed.numel() != pkb or self._packed.device != device:
                self._packed = torch.empty(pkb, dtype=torch.uint8, device=device)
            check(L.b200fno_plan_bind(plan, self._ws.data_ptr(), wsb, self._packed.data_ptr(), pkb))
        self._max_batch, self._device, self._weights_key = batch, device, None

    def set_compute(self, compute: str) -> None:
        """'f32': fp32 semantics (3xTF32 on the tensor cores).  'bf16': the reference under
        ``torch.autocast(dtype=torch.bfloat16)`` - Linear / Conv operands cast to bf16, one tensor-core
        pass with fp32 accumulation; FFT stages, mode mixing, BatchNorm and the stored tensors stay fp32."""
        if compute not in _capi.COMPUTE:
            raise ValueError(f"compute must be one of {sorted(_capi.COMPUTE)}, got {compute!r}")
        if compute == self.compute:
            return
        self.compute = compute
        if self._plan is not None:
            check(_capi.lib().b200fno_plan_set_compute(self._plan, _capi.COMPUTE[compute]))
            self._weights_key = None  # the packed copies depend on the mode

    # -- weights --------------------------------------------------------------
    def _pack(self, sd: dict, stream: int):
        """sd: reference-layout tensors on the plan's device (see fno.py: module.engine_state())."""
        L = _capi.lib()
        ncorner = 4 if self.ndim == 3 else 2
        keep = []

        def f32(t):
            t = t.detach()
            if t.is_complex():
                t = torch.view_as_real(t)
            t = t.to(device=self._device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        spec, conv_w, conv_b, bw, bb, bm, bv = [], [], [], [], [], [], []
        for i in range(self.n_layers):
            for k in range(ncorner):
                spec.append(f32(sd[f"spectral_convs.{i}.weights{k + 1}"]))
            conv_w.append(f32(sd[f"convs.{i}.weight"]))
            conv_b.append(f32(sd[f"convs.{i}.bias"]))
            bw.append(f32(sd[f"bns.{i}.weight"]))
            bb.append(f32(sd[f"bns.{i}.bias"]))
            bm.append(f32(sd[f"bns.{i}.running_mean"]))
            bv.append(f32(sd[f"bns.{i}.running_var"]))
        arrs = [_capi.ptr_array(a) for a in (spec, conv_w, conv_b, bw, bb, bm, bv)]
        w = Weights(fc0_w=f32(sd["fc0.weight"]), fc0_b=f32(sd["fc0.bias"]), spec_w=arrs[0], conv_w=arrs[1],
                    conv_b=arrs[2], bn_weight=arrs[3], bn_bias=arrs[4], bn_mean=arrs[5], bn_var=arrs[6],
                    fc1_w=f32(sd["fc1.weight"]), fc1_b=f32(sd["fc1.bias"]), fc2_w=f32(sd["fc2.weight"]),
                    fc2_b=f32(sd["fc2.bias"]))
        check(L.b200fno_pack_weights(self._plan, C.byref(w), stream))
        self._keepalive = keep  # sources must outlive the enqueued pack kernels

    def prepare(self, batch: int, device: torch.device, sd: dict, key) -> int:
        """Make the plan fit ``batch`` on ``device`` and the packed weights current.  Returns the stream handle."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("b200fno: the engine runs on CUDA (sm_100a) only and has no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._ensure_plan(batch, device)
        stream = torch.cuda.current_stream(device).cuda_stream
        if key != self._weights_key:
            with torch.cuda.device(device):
                self._pack(sd, stream)
            self._weights_key = key
        return stream

    # -- hot path ---------------------------------------------------------------
    def forward(self, x: torch.Tensor, sd: dict, key) -> torch.Tensor:
        _require_cuda(x, "input")
        if tuple(x.shape[1:]) != self.shape_in:
            raise RuntimeError(f"b200fno: input shape {tuple(x.shape)} does not match [B,{self.shape_in}]")
        x = x.contiguous()
        stream = self.prepare(x.shape[0], x.device, sd, key)
        y = torch.empty((x.shape[0], *self.shape_out), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_capi.lib().b200fno_forward(self._plan, x.shape[0], x.data_ptr(), y.data_ptr(), stream))
        return y

    def rollout(self, x0: torch.Tensor, a: torch.Tensor, b: torch.Tensor, n_steps: int, sd: dict, key,
                out: Optional[torch.Tensor] = None, graph: bool = False) -> torch.Tensor:
        """x0: normalised input [B,T,H,W,C_in]; a, b: per-output-channel affine (len C_out).
        Returns cat(preds[1:], dim=1)[..., :C_out] of eval.py:313-322, shape [B, n*T_out, H, W, C_out].

        ``graph=True``: the ~22-30 launches per step are captured ONCE per (batch, n_steps) into a CUDA graph and
        replayed (small batches are launch-latency bound).  The graph owns static input / affine / output buffers:
        the returned tensor is that static output (valid until the next graph rollout of the same shape) unless
        ``out`` is given, in which case the result is copied into it."""
        if graph:
            return self._rollout_graph(x0, a, b, n_steps, sd, key, out)
        _require_cuda(x0, "input")
        if tuple(x0.shape[1:]) != self.shape_in:
            raise RuntimeError(f"b200fno: input shape {tuple(x0.shape)} does not match [B,{self.shape_in}]")
        x0 = x0.contiguous()
        B = x0.shape[0]
        stream = self.prepare(B, x0.device, sd, key)
        t_out, h, w, c_out = self.shape_out
        a = a.to(device=x0.device, dtype=torch.float32).contiguous()
        b = b.to(device=x0.device, dtype=torch.float32).contiguous()
        if a.numel() != c_out or b.numel() != c_out:
            raise RuntimeError("b200fno: affine vectors must have C_out entries")
        if out is None:
            out = torch.empty((B, n_steps * t_out, h, w, c_out), dtype=torch.float32, device=x0.device)
        # c_in == c_out: the engine feeds each step from the prediction slice the previous one wrote (no state buffer)
        need_state = n_steps > 1 and self.shape_in[3] != self.shape_out[3]
        state = torch.empty((2, *x0.shape), dtype=torch.float32, device=x0.device) if need_state else None
        with torch.cuda.device(x0.device):
            check(_capi.lib().b200fno_rollout(self._plan, B, x0.data_ptr(), a.data_ptr(), b.data_ptr(), n_steps,
                                              state.data_ptr() if state is not None else None, out.data_ptr(),
                                              stream))
        return out

    def _rollout_graph(self, x0, a, b, n_steps, sd, key, out):
        _require_cuda(x0, "input")
        B = x0.shape[0]
        self.prepare(B, x0.device, sd, key)  # plan + packed weights OUTSIDE the capture (a re-pack keeps its addresses)
        gkey = (B, n_steps, self.compute, id(self._plan))
        graphs = self.__dict__.setdefault("_graphs", {})
        g = graphs.get(gkey)
        if g is None:
            t_out, h, w, c_out = self.shape_out
            st = {"x": torch.empty_like(x0, memory_format=torch.contiguous_format),
                  "a": torch.empty(c_out, dtype=torch.float32, device=x0.device),
                  "b": torch.empty(c_out, dtype=torch.float32, device=x0.device),
                  "out": torch.empty((B, n_steps * t_out, h, w, c_out), dtype=torch.float32, device=x0.device)}
            st["x"].copy_(x0), st["a"].copy_(a), st["b"].copy_(b)
            side = torch.cuda.Stream(device=x0.device)
            side.wait_stream(torch.cuda.current_stream(x0.device))
            with torch.cuda.stream(side):  # warm-up on the capture stream (lazy module loading must not be captured)
                self.rollout(st["x"], st["a"], st["b"], n_steps, sd, key, out=st["out"])
            torch.cuda.current_stream(x0.device).wait_stream(side)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                # the state ping-pong buffer is allocated inside the capture: it lives in the graph's private pool
                self.rollout(st["x"], st["a"], st["b"], n_steps, sd, key, out=st["out"])
            g = graphs[gkey] = (cg, st)
        cg, st = g
        st["x"].copy_(x0), st["a"].copy_(a.to(st["a"].device, torch.float32)), st["b"].copy_(b.to(st["b"].device, torch.float32))
        cg.replay()
        if out is not None:
            out.copy_(st["out"])
            return out
        return st["out"]

    # -- training path ----------------------------------------------------------
    def _ensure_train_ws(self, device: torch.device):
        if self._train_ws is not None:
            return
        L = _capi.lib()
        nbytes = L.b200fno_train_workspace_bytes(self._plan)
        self._train_ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if os.environ.get("B200FNO_POISON_WS"):
            self._train_ws.fill_(0xFF)
        with torch.cuda.device(device):
            check(L.b200fno_train_bind(self._plan, self._train_ws.data_ptr(), nbytes))

    def train_forward(self, x: torch.Tensor, sd: dict, key, running_mean: Sequence[Optional[torch.Tensor]],
                      running_var: Sequence[Optional[torch.Tensor]], momentum: float) -> torch.Tensor:
        """FNO3d.forward in .train() mode (batch-statistics BatchNorm, running buffers updated in place);
        keeps what ``train_backward`` needs inside the engine's training workspace."""
        _require_cuda(x, "input")
        if tuple(x.shape[1:]) != self.shape_in:
            raise RuntimeError(f"b200fno: input shape {tuple(x.shape)} does not match [B,{self.shape_in}]")
        x = x.contiguous()
        stream = self.prepare(x.shape[0], x.device, sd, key)
        self._ensure_train_ws(x.device)
        for t in list(running_mean) + list(running_var):
            if t is not None:
                _require_cuda(t, "BatchNorm running statistic")
        rm = _capi.ptr_array([t.data_ptr() for t in running_mean]) if all(t is not None for t in running_mean) else None
        rv = _capi.ptr_array([t.data_ptr() for t in running_var]) if all(t is not None for t in running_var) else None
        y = torch.empty((x.shape[0], *self.shape_out), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_capi.lib().b200fno_train_forward(self._plan, x.shape[0], x.data_ptr(), y.data_ptr(), rm, rv,
                                                    float(momentum), stream))
        self._train_seq += 1
        return y

    @property
    def train_seq(self) -> int:
        return self._train_seq

    def train_backward(self, x: torch.Tensor, dy: torch.Tensor, params: dict,
                       ready_events: Optional[Sequence[int]] = None, seq: Optional[int] = None,
                       input_grad: bool = False) -> dict:
        """Parameter gradients of the last ``train_forward`` in the reference layout.
        ``params``: name -> parameter tensor (reference state_dict names); returns name -> gradient tensor.
        ``ready_events``: optional ``n_layers + 1`` raw ``cudaEvent_t`` handles recorded when each gradient group
        is final (see ``b200fno_train_backward``).  ``input_grad``: also return the gradient with respect to the
        input field under the key ``"__input__"`` (same layout as ``x``)."""
        _require_cuda(dy, "output gradient")
        if seq is not None and seq != self._train_seq:
            # the training workspace holds the activations of ONE forward (b200fno.h: "one forward outstanding per
            # plan"); a second train-mode forward before this backward has overwritten them
            raise RuntimeError(
                f"b200fno: backward of train-mode forward #{seq} requested, but the engine's training workspace holds "
                f"forward #{self._train_seq}: only one train-mode forward may be outstanding per model (run "
                "forward -> backward pairs, e.g. accumulate gradients over separate loss.backward() calls)")
        x, dy = x.contiguous(), dy.contiguous()
        ncorner = 4 if self.ndim == 3 else 2
        if os.environ.get("B200FNO_POISON_GRADS"):  # debugging aid: a never-written gradient element stays NaN
            grads = {k: torch.full_like(v, float("nan")) for k, v in params.items()}
        else:
            grads = {k: torch.empty_like(v) for k, v in params.items()}

        def ptr(name):
            g = grads[name]
            return (torch.view_as_real(g) if g.is_complex() else g).data_ptr()

        n = self.n_layers
        arrs = [
            _capi.ptr_array([ptr(f"spectral_convs.{i}.weights{k + 1}") for i in range(n) for k in range(ncorner)]),
            _capi.ptr_array([ptr(f"convs.{i}.weight") for i in range(n)]),
            _capi.ptr_array([ptr(f"convs.{i}.bias") for i in range(n)]),
            _capi.ptr_array([ptr(f"bns.{i}.weight") for i in range(n)]),
            _capi.ptr_array([ptr(f"bns.{i}.bias") for i in range(n)]),
        ]
        dx = torch.empty_like(x) if input_grad else None
        g = Grads(fc0_w=ptr("fc0.weight"), fc0_b=ptr("fc0.bias"), spec_w=arrs[0], conv_w=arrs[1], conv_b=arrs[2],
                  bn_weight=arrs[3], bn_bias=arrs[4], fc1_w=ptr("fc1.weight"), fc1_b=ptr("fc1.bias"),
                  fc2_w=ptr("fc2.weight"), fc2_b=ptr("fc2.bias"), x=dx.data_ptr() if input_grad else None)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            ev = _capi.ptr_array(list(ready_events)) if ready_events is not None else None
            check(_capi.lib().b200fno_train_backward(self._plan, x.shape[0], x.data_ptr(), dy.data_ptr(), C.byref(g),
                                                     ev, stream))
        if input_grad:
            grads["__input__"] = dx
        return grads

    def resolved_impl(self) -> str:
        """'tc' if the tcgen05 layer kernel is in use for this shape, else 'simt' (plan must exist)."""
        return "tc" if _capi.lib().b200fno_plan_get_impl(self._plan) == _capi.IMPL_TC else "simt"

    def stage_impls(self) -> dict:
        """stage name -> 'tc' | 'simt' for the bound plan (b200fno_plan_stage_impl)."""
        out = {}
        for i, s in enumerate(_capi.STAGES):
            r = _capi.lib().b200fno_plan_stage_impl(self._plan, i)
            if r < 0:
                check(r)
            out[s] = "tc" if r else "simt"
        return out

    def timing(self, on: bool) -> None:
        """Bracket every stage launch with CUDA events (bench.py's per-kernel roofline numbers)."""
        check(_capi.lib().b200fno_timing_enable(self._plan, int(on)))

    def timing_collect(self) -> dict:
        ms = (C.c_double * len(_capi.STAGES))()
        cnt = (C.c_int64 * len(_capi.STAGES))()
        check(_capi.lib().b200fno_timing_collect(self._plan, ms, cnt))
        return {s: {"ms": ms[i], "launches": cnt[i]} for i, s in enumerate(_capi.STAGES) if cnt[i]}

    def algorithmic_bytes(self, batch: int) -> float:
        return float(_capi.lib().b200fno_algorithmic_bytes(self._plan, batch))


def spectral_conv(x: torch.Tensor, weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """Stand-alone SpectralConv{2,3}d.forward (fno.py:45-64) in the reference (channels-first) layout."""
    _require_cuda(x, "input")
    ndim = x.dim() - 2
    if ndim not in (2, 3):
        raise RuntimeError("b200fno: spectral_conv expects [B,C,H,W] or [B,C,T,H,W]")
    L = _capi.lib()
    x = x.contiguous()
    ci, co = weights[0].shape[0], weights[0].shape[1]
    if x.shape[1] != ci:
        raise RuntimeError(f"b200fno: input has {x.shape[1]} channels, weights expect {ci}")
    m = tuple(weights[0].shape[2:])
    m1, m2, m3 = (m if ndim == 3 else (1, *m))
    t, h, w = (x.shape[2:] if ndim == 3 else (1, *x.shape[2:]))
    B = x.shape[0]
    wr = [torch.view_as_real(wt.detach().to(device=x.device, dtype=torch.complex64).contiguous()) for wt in weights]
    need = L.b200fno_spectral_workspace_bytes(ndim, B, ci, co, t, h, w, m1, m2, m3)
    if need == 0:
        check(-1)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    y = torch.empty((B, co, *x.shape[2:]), dtype=torch.float32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        check(L.b200fno_spectral_conv(ndim, B, ci, co, t, h, w, m1, m2, m3,
                                      _capi.ptr_array([t_.data_ptr() for t_ in wr]), x.data_ptr(), y.data_ptr(),
                                      ws.data_ptr(), need, stream))
    return y
