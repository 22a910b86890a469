"""CPU oracle for the FNO forward / rollout hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a ``state_dict``-shaped mapping of
tensors, the algorithm of

* ``realpdebench/model/fno.py:16-64``   (SpectralConv3d)
* ``realpdebench/model/fno.py:66-143``  (FNO3d incl. ``get_grid``)
* ``realpdebench/data/data_normalizer.py:6-62,98-130`` (the three normalisers)
* ``realpdebench/eval.py:305-326``      (the autoregressive rollout loop)
* ``realpdebench/train.py:321-334``     (one optimisation step; gradients by torch autograd of the restated forward)

The arithmetic of that path lives in PyTorch/ATen (``torch.fft.rfftn/irfftn``,
``einsum``, ``conv3d``, ``batch_norm``, ``gelu``, ``linear``; torch is an
un-pinned dependency of the reference, ``pyproject.toml:39``), so the
restatement calls the same torch CPU operators in the same order: on the same
torch build it is bit-identical to the reference modules.

Parity pin: the reference ships no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference itself, executed in the build container from ``/root/reference`` by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.pt``, plus
the two known-answer vectors KAT-A / KAT-B recorded in SURVEY.md section 4.
``tests/test_oracle.py`` checks both.

FNO-2D does not exist in the reference (``model/fno.py`` is 3-D only).  The
2-D functions here are the frozen definition from SURVEY.md section 8(c): the
3-D file with the time axis folded into channels.  Its spectral layer is
pinned against ``realpdebench/model/MWT_libs/models.py:252-289``
(``sparseKernelFT2d`` spectral math) in the golden set.

All functions accept ``dtype=torch.float64`` inputs/weights as well, which the
tests use to decide which of two fp32 results is closer to the truth.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
PADDING = 6  # fno.py:87
PROJ_HIDDEN = 128  # fno.py:102


# --------------------------------------------------------------------------
# spectral layers
# --------------------------------------------------------------------------
def spectral_conv3d(x: Tensor, w1: Tensor, w2: Tensor, w3: Tensor, w4: Tensor) -> Tensor:
    """fno.py:45-64.  x [B,Ci,T,H,W] real; w* [Ci,Co,m1,m2,m3] complex."""
    m1, m2, m3 = w1.shape[2:]
    b, co = x.shape[0], w1.shape[1]
    x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])  # fno.py:48
    out_ft = torch.zeros(b, co, x.size(-3), x.size(-2), x.size(-1) // 2 + 1,
                         dtype=x_ft.dtype, device=x.device)  # fno.py:51
    mul = lambda a, w: torch.einsum("bixyz,ioxyz->boxyz", a, w)  # fno.py:41-43
    out_ft[:, :, :m1, :m2, :m3] = mul(x_ft[:, :, :m1, :m2, :m3], w1)      # fno.py:53
    out_ft[:, :, -m1:, :m2, :m3] = mul(x_ft[:, :, -m1:, :m2, :m3], w2)    # fno.py:55
    out_ft[:, :, :m1, -m2:, :m3] = mul(x_ft[:, :, :m1, -m2:, :m3], w3)    # fno.py:57
    out_ft[:, :, -m1:, -m2:, :m3] = mul(x_ft[:, :, -m1:, -m2:, :m3], w4)  # fno.py:59
    return torch.fft.irfftn(out_ft, s=(x.size(-3), x.size(-2), x.size(-1)))  # fno.py:63


def spectral_conv2d(x: Tensor, w1: Tensor, w2: Tensor) -> Tensor:
    """2-D analogue (MWT_libs/models.py:270-289).  x [B,Ci,H,W]."""
    m1, m2 = w1.shape[2:]
    b, co = x.shape[0], w1.shape[1]
    x_ft = torch.fft.rfft2(x)
    out_ft = torch.zeros(b, co, x.size(-2), x.size(-1) // 2 + 1, dtype=x_ft.dtype, device=x.device)
    mul = lambda a, w: torch.einsum("bixy,ioxy->boxy", a, w)
    out_ft[:, :, :m1, :m2] = mul(x_ft[:, :, :m1, :m2], w1)
    out_ft[:, :, -m1:, :m2] = mul(x_ft[:, :, -m1:, :m2], w2)
    return torch.fft.irfft2(out_ft, s=(x.size(-2), x.size(-1)))


# --------------------------------------------------------------------------
# whole networks
# --------------------------------------------------------------------------
def _linspace(n: int, dtype) -> Tensor:
    # fno.py:137 — NumPy float64 linspace, then cast
    return torch.tensor(np.linspace(0, 1, n), dtype=dtype)


def grid3d(shape: Sequence[int], dtype=torch.float) -> Tensor:
    """fno.py:135-143.  shape = (B,T,H,W,...) -> [B,T,H,W,3]."""
    b, sx, sy, sz = shape[0], shape[1], shape[2], shape[3]
    gx = _linspace(sx, dtype).reshape(1, sx, 1, 1, 1).repeat([b, 1, sy, sz, 1])
    gy = _linspace(sy, dtype).reshape(1, 1, sy, 1, 1).repeat([b, sx, 1, sz, 1])
    gz = _linspace(sz, dtype).reshape(1, 1, 1, sz, 1).repeat([b, sx, sy, 1, 1])
    return torch.cat((gx, gy, gz), dim=-1)


def grid2d(b: int, h: int, w: int, dtype=torch.float) -> Tensor:
    gx = _linspace(h, dtype).reshape(1, h, 1, 1).repeat([b, 1, w, 1])
    gy = _linspace(w, dtype).reshape(1, 1, w, 1).repeat([b, h, 1, 1])
    return torch.cat((gx, gy), dim=-1)


def _bn(x: Tensor, sd: Mapping[str, Tensor], i: int, training: bool) -> Tensor:
    p = f"bns.{i}."
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], training, 0.1, 1e-5)


def n_layers_of(sd: Mapping[str, Tensor]) -> int:
    n = 0
    while f"convs.{n}.weight" in sd:
        n += 1
    return n


def fno3d_forward(sd: Mapping[str, Tensor], x: Tensor, shape_out: Sequence[int],
                  training: bool = False, padding: int = PADDING) -> Tensor:
    """fno.py:105-129.  x [B,T,H,W,C_in] -> [B,T_out,H,W,C_out].

    ``sd`` uses the reference ``state_dict`` keys.  With ``training=True`` the
    BatchNorm layers use batch statistics and update the running buffers in
    ``sd`` in place, exactly like the reference module in ``.train()`` mode.
    """
    t_in = x.shape[1]
    r = shape_out[0] // t_in
    grid = grid3d(x.shape, x.dtype).to(x.device)
    h = torch.cat((x, grid), dim=-1)
    h = F.linear(h, sd["fc0.weight"], sd["fc0.bias"])
    h = h.permute(0, 4, 1, 2, 3)
    h = F.pad(h, [0, padding, 0, padding, 0, padding])
    L = n_layers_of(sd)
    for i in range(L):
        p = f"spectral_convs.{i}."
        x1 = spectral_conv3d(h, sd[p + "weights1"], sd[p + "weights2"], sd[p + "weights3"], sd[p + "weights4"])
        x2 = F.conv3d(h, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"])
        h = x1 + x2
        h = _bn(h, sd, i, training)
        if i < L - 1:
            h = F.gelu(h)
    h = h[..., :-padding, :-padding, :-padding]
    h = h.permute(0, 2, 3, 4, 1)
    h = F.linear(h, sd["fc1.weight"], sd["fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd["fc2.weight"], sd["fc2.bias"])
    h = h.reshape(*h.shape[:-1], shape_out[-1], r)  # fno.py:127
    return h.permute(0, 1, 5, 2, 3, 4).reshape(h.shape[0], *shape_out)  # fno.py:128


def fno2d_forward(sd: Mapping[str, Tensor], x: Tensor, shape_out: Sequence[int],
                  training: bool = False, padding: int = PADDING) -> Tensor:
    """FNO-2D: frames folded into channels.  x [B,T,H,W,C_in].

    lift feature j = t*C_in + c (then the two grid coordinates h, w);
    projection feature f = t_out*C_out + c.
    """
    b, t, hh, ww, ci = x.shape
    t_out, _, _, co = shape_out
    h = x.permute(0, 2, 3, 1, 4).reshape(b, hh, ww, t * ci)
    h = torch.cat((h, grid2d(b, hh, ww, x.dtype).to(x.device)), dim=-1)
    h = F.linear(h, sd["fc0.weight"], sd["fc0.bias"])
    h = h.permute(0, 3, 1, 2)
    h = F.pad(h, [0, padding, 0, padding])
    L = n_layers_of(sd)
    for i in range(L):
        p = f"spectral_convs.{i}."
        x1 = spectral_conv2d(h, sd[p + "weights1"], sd[p + "weights2"])
        x2 = F.conv2d(h, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"])
        h = x1 + x2
        h = _bn(h, sd, i, training)
        if i < L - 1:
            h = F.gelu(h)
    h = h[..., :-padding, :-padding]
    h = h.permute(0, 2, 3, 1)
    h = F.linear(h, sd["fc1.weight"], sd["fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd["fc2.weight"], sd["fc2.bias"])  # [B,H,W,T_out*C_out]
    return h.reshape(b, hh, ww, t_out, co).permute(0, 3, 1, 2, 4).contiguous()


# --------------------------------------------------------------------------
# parameter initialisation in the reference's RNG order
# --------------------------------------------------------------------------
def _linear_init(out_f: int, in_f: int) -> Tuple[Tensor, Tensor]:
    """torch.nn.Linear.reset_parameters: kaiming_uniform(a=sqrt(5)) then bias."""
    w = torch.empty(out_f, in_f)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1 / math.sqrt(in_f) if in_f > 0 else 0
    b = torch.empty(out_f)
    torch.nn.init.uniform_(b, -bound, bound)
    return w, b


def init_state(ndim: int, modes: Sequence[int], n_layers: int, width: int,
               shape_in: Sequence[int], shape_out: Sequence[int]) -> Dict[str, Tensor]:
    """Draw parameters in the order fno.py:89-103 constructs them.

    With the same ``torch.manual_seed`` this reproduces the reference FNO3d
    weights bit for bit (ndim=3).  ndim=2 follows the same order with 2 corner
    weights per layer.
    """
    sd: Dict[str, Tensor] = {}
    t_in, c_in = shape_in[0], shape_in[-1]
    t_out, c_out = shape_out[0], shape_out[-1]
    if ndim == 3:
        dim_in, dim_out = c_in + 3, c_out * t_out // t_in
    else:
        dim_in, dim_out = t_in * c_in + 2, t_out * c_out
    sd["fc0.weight"], sd["fc0.bias"] = _linear_init(width, dim_in)
    scale = 1 / (width * width)
    ncorner = 4 if ndim == 3 else 2
    for i in range(n_layers):
        for k in range(ncorner):
            sd[f"spectral_convs.{i}.weights{k + 1}"] = scale * torch.rand(width, width, *modes, dtype=torch.cfloat)
        w, b = _linear_init(width, width)  # Conv*d(k=1) init == Linear init with fan_in=width
        sd[f"convs.{i}.weight"] = w.reshape(width, width, *([1] * ndim))
        sd[f"convs.{i}.bias"] = b
        sd[f"bns.{i}.weight"] = torch.ones(width)
        sd[f"bns.{i}.bias"] = torch.zeros(width)
        sd[f"bns.{i}.running_mean"] = torch.zeros(width)
        sd[f"bns.{i}.running_var"] = torch.ones(width)
        sd[f"bns.{i}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["fc1.weight"], sd["fc1.bias"] = _linear_init(PROJ_HIDDEN, width)
    sd["fc2.weight"], sd["fc2.bias"] = _linear_init(dim_out, PROJ_HIDDEN)
    return sd


def randomize_bn(sd: Dict[str, Tensor], seed: int = 123) -> None:
    """KAT-A recipe: non-trivial BN statistics, in layer order."""
    g = torch.Generator().manual_seed(seed)
    for i in range(n_layers_of(sd)):
        c = sd[f"bns.{i}.weight"].numel()
        sd[f"bns.{i}.running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[f"bns.{i}.running_var"] = torch.rand(c, generator=g) + 0.5
        sd[f"bns.{i}.weight"] = torch.rand(c, generator=g) + 0.5
        sd[f"bns.{i}.bias"] = torch.randn(c, generator=g) * 0.1


# --------------------------------------------------------------------------
# normalisers (data_normalizer.py) and rollout (eval.py:305-326)
# --------------------------------------------------------------------------
class Normalizer:
    """kind in {'none','gaussian','range'}; stats are 1-D per-channel tensors."""

    def __init__(self, kind: str, device="cpu", mean_inputs=None, std_inputs=None,
                 mean_targets=None, std_targets=None, max_inputs=None, max_targets=None):
        self.kind, self.device = kind, device
        fix = lambda s: torch.where(s == 0, torch.ones_like(s), s).to(device)  # data_normalizer.py:47-48
        if kind == "gaussian":
            self.mean_inputs, self.mean_targets = mean_inputs.to(device), mean_targets.to(device)
            self.std_inputs, self.std_targets = fix(std_inputs), fix(std_targets)
        elif kind == "range":
            self.max_inputs, self.max_targets = fix(max_inputs), fix(max_targets)
        elif kind != "none":
            raise ValueError(f"Normalizer {kind} not supported")  # eval.py:275

    def preprocess(self, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
        c1, c2 = x.shape[-1], y.shape[-1]
        x, y = x.to(self.device), y.to(self.device)
        if self.kind == "gaussian":  # data_normalizer.py:50-55
            x = (x - self.mean_inputs[..., :c1]) / self.std_inputs[..., :c1]
            y = (y - self.mean_targets[..., :c2]) / self.std_targets[..., :c2]
        elif self.kind == "range":  # data_normalizer.py:118-123
            x = x / self.max_inputs[..., :c1]
            y = y / self.max_targets[..., :c2]
        return x, y

    def postprocess(self, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
        c1, c2 = x.shape[-1], y.shape[-1]
        x, y = x.to(self.device), y.to(self.device)
        if self.kind == "gaussian":  # data_normalizer.py:57-62
            x = x * self.std_inputs[..., :c1] + self.mean_inputs[..., :c1]
            y = y * self.std_targets[..., :c2] + self.mean_targets[..., :c2]
        elif self.kind == "range":  # data_normalizer.py:125-130
            x = x * self.max_inputs[..., :c1]
            y = y * self.max_targets[..., :c2]
        return x, y


def synthetic_normalizer(c_in: int, c_out: int, seed: int = 4321, kind: str = "gaussian") -> Normalizer:
    """mu ~ N(0,1)*0.1, sigma ~ U(0.5,1.5) from Generator(seed)."""
    g = torch.Generator().manual_seed(seed)
    mi, si = torch.randn(c_in, generator=g) * 0.1, torch.rand(c_in, generator=g) + 0.5
    mt, st = torch.randn(c_out, generator=g) * 0.1, torch.rand(c_out, generator=g) + 0.5
    if kind == "gaussian":
        return Normalizer(kind, mean_inputs=mi, std_inputs=si, mean_targets=mt, std_targets=st)
    if kind == "range":
        return Normalizer(kind, max_inputs=si * 3, max_targets=st * 3)
    return Normalizer("none")


def rollout(model_fn: Callable[[Tensor], Tensor], norm: Normalizer, input: Tensor, target: Tensor,
            n_autoregressive: int, teacher: Optional[Sequence[Tensor]] = None):
    """eval.py:296-326 for one batch.

    Returns ``(pred, target, normalized_loss, states)`` where pred/target are the
    de-normalised tensors eval.py appends to its lists (:342-343),
    ``normalized_loss`` is the scalar added to ``normalized_test_loss`` (:323)
    and ``states`` is the list ``preds`` (normalised model inputs per step).

    ``teacher``: optional list of normalised states; if given, step i is fed
    ``teacher[i]`` instead of the previous prediction (per-step parity checks).
    """
    b = input.size(0)
    unmeasured_c = sum(int(torch.all(target[..., c_] == 0)) for c_ in range(target.shape[-1]))  # :298-302
    c = target.shape[-1] - unmeasured_c
    in_control = input.shape[-1] != target.shape[-1]  # :305-309
    if in_control:
        para_c = input.shape[-1] - target.shape[-1]
        para_input = input[..., -para_c:]
    input, target = norm.preprocess(input, target)  # :311
    preds = [input]
    for i in range(n_autoregressive):  # :313-319
        p = model_fn(preds[-1] if teacher is None else teacher[i])
        _, p = norm.postprocess(preds[-1], p)
        if in_control:
            p = torch.cat([p, para_input.to(p.device)], dim=-1)
        p, _ = norm.preprocess(p, target)
        preds.append(p)
    pred = torch.cat(preds[1:], dim=1)  # :321
    if in_control:
        pred = pred[..., :-para_c]
    loss = F.mse_loss(pred[..., :c], target[..., :c], reduction="none").reshape(b, -1).mean().item()  # :323
    _, pred = norm.postprocess(input, pred)  # :325
    _, target = norm.postprocess(input, target)  # :326
    return pred, target, loss, preds


# --------------------------------------------------------------------------
# training step (train.py:321-334)
# --------------------------------------------------------------------------
PARAM_SUFFIXES = ("weight", "bias", "weights1", "weights2", "weights3", "weights4")


def is_param(name: str) -> bool:
    """state_dict entries that are nn.Parameters (everything but the BatchNorm buffers)."""
    return name.rsplit(".", 1)[-1] in PARAM_SUFFIXES


def train_loss_and_grads(ndim: int, sd: Dict[str, Tensor], x: Tensor, target: Tensor, shape_out: Sequence[int],
                         input_grad: bool = False):
    """``loss = model.train_loss(input, target).mean(); loss.backward()`` (train.py:328-329, fno.py:131-133).

    Returns ``(loss, grads, pred)``; ``sd``'s BatchNorm running buffers are updated in place like the
    reference module in ``.train()`` mode (``num_batches_tracked`` included).  ``input_grad``: ``grads["__input__"]``
    = the gradient with respect to ``x`` (autograd through the same forward)."""
    leaves = {k: (v.detach().clone().requires_grad_(True) if is_param(k) else v) for k, v in sd.items()}
    fwd = fno3d_forward if ndim == 3 else fno2d_forward
    if input_grad:
        x = x.detach().clone().requires_grad_(True)
    pred = fwd(leaves, x, shape_out, training=True)
    loss = F.mse_loss(pred, target, reduction="none").mean()  # utils/metrics.py:11-13 + train.py:328
    loss.backward()
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k] = sd[k] + 1
    grads = {k: v.grad for k, v in leaves.items() if is_param(k)}
    if input_grad:
        grads["__input__"] = x.grad
    return float(loss.detach()), grads, pred.detach()


def train_steps(ndim: int, sd: Dict[str, Tensor], batches, shape_out: Sequence[int], lr: float,
                clip_grad_norm: float = 0.0, step_size: int = 100):
    """train.py:321-334 for a list of normalised ``(input, target)`` batches: Adam + StepLR(gamma=0.5).
    Updates ``sd`` in place; returns the list of losses."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if is_param(k)}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=0.5)
    fwd = fno3d_forward if ndim == 3 else fno2d_forward
    losses = []
    state = dict(sd)
    for x, target in batches:
        opt.zero_grad()
        state.update(params)
        loss = F.mse_loss(fwd(state, x, shape_out, training=True), target, reduction="none").mean()
        loss.backward()
        if clip_grad_norm > 0:
            torch.nn.utils.clip_grad_norm_(list(params.values()), clip_grad_norm)
        opt.step()
        sched.step()
        losses.append(float(loss.detach()))
    for k, v in params.items():
        sd[k] = v.detach()
    for k in sd:
        if not is_param(k) and not k.endswith("num_batches_tracked"):
            sd[k] = state[k]
        elif k.endswith("num_batches_tracked"):
            sd[k] = sd[k] + len(losses)
    return losses


def rel_l2(a: Tensor, b: Tensor) -> float:
    """||a-b||_2 / ||b||_2 over the whole tensor, in float64."""
    a, b = (torch.view_as_real(t.detach()) if t.is_complex() else t.detach() for t in (a, b))
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
