"""Drop-in for ``realpdebench.model.fno`` (reference realpdebench/model/fno.py).

Same class names, constructor signatures, parameter names / shapes / dtypes and
initialisation order as the reference, so ``state_dict()`` / ``load_state_dict``
/ ``parameters()`` and released checkpoints are interchangeable
(checkpoint row).  ``forward`` runs on the CUDA engine.

The parameters stay in the reference layout (that is what optimisers and
checkpoints see); the engine keeps its own packed copy and refreshes it
whenever a parameter or BatchNorm buffer changes (tensor version counters).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import FNOEngine, spectral_conv
from .model import Model


def mse_loss(pred, target):
    """realpdebench/utils/metrics.py:11-13 — unreduced squared error."""
    return nn.functional.mse_loss(pred, target, reduction='none')


class SpectralConv3d(nn.Module):
    """fno.py:16-64.  weights1..4: complex64 [Ci,Co,m1,m2,m3], init scale*rand."""

    def __init__(self, in_channels, out_channels, modes1, modes2, modes3):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2, self.modes3 = modes1, modes2, modes3
        self.scale = 1 / (in_channels * out_channels)
        shape = (in_channels, out_channels, modes1, modes2, modes3)
        for k in (1, 2, 3, 4):  # same RNG order as fno.py:31-38
            setattr(self, f"weights{k}", nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.cfloat)))

    def forward(self, x):
        return spectral_conv(x, [self.weights1, self.weights2, self.weights3, self.weights4])


class SpectralConv2d(nn.Module):
    """2-D analogue: weights1 (low H), weights2 (high H): complex64 [Ci,Co,m1,m2]."""

    def __init__(self, in_channels, out_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1 / (in_channels * out_channels)
        for k in (1, 2):
            setattr(self, f"weights{k}",
                    nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, modes2, dtype=torch.cfloat)))

    def forward(self, x):
        return spectral_conv(x, [self.weights1, self.weights2])


class _TrainForward(torch.autograd.Function):
    """Train-mode forward / backward on the engine (b200fno_train_forward / b200fno_train_backward).

    The parameters are inputs of the Function, so autograd delivers the engine's gradients to the very
    ``nn.Parameter`` objects ``train.py:290`` hands to Adam (and DDP-style hooks fire).  The gradient with
    respect to the input field is produced when the input requires it (a caller differentiating through the
    surrogate; the reference training loop itself never does)."""

    @staticmethod
    def forward(ctx, module, x, names, *params):
        sd, key = module._engine_state()
        bns = list(module.bns)
        track = all(bn.track_running_stats and bn.running_mean is not None for bn in bns)
        rm = [bn.running_mean if track else None for bn in bns]
        rv = [bn.running_var if track else None for bn in bns]
        if any(bn.momentum is None for bn in bns):
            raise NotImplementedError("b200fno: BatchNorm momentum=None (cumulative average) is not supported; the "
                                      "reference uses the default momentum 0.1 (fno.py:100)")
        momentum = bns[0].momentum
        y = module._engine.train_forward(x, sd, key, rm, rv, momentum)
        if track:
            for bn in bns:
                bn.num_batches_tracked += 1  # nn.BatchNorm bookkeeping (buffer, also in state_dict)
            # the engine updated the running buffers in place behind torch's version counters
            module._stats_epoch = getattr(module, "_stats_epoch", 0) + 1
        ctx.module, ctx.names, ctx.x = module, names, x
        ctx.params = params
        ctx.seq = module._engine.train_seq  # the workspace now holds THIS forward's activations
        return y

    @staticmethod
    def backward(ctx, dy):
        need = ctx.needs_input_grad[3:]
        params = dict(zip(ctx.names, ctx.params))
        sync = getattr(ctx.module, "_grad_sync", None)  # dist.OverlappedGradientReducer: all-reduce under the backward
        want_dx = bool(ctx.needs_input_grad[1])
        if sync is not None:
            grads = sync.backward_and_reduce(ctx.module._engine, ctx.x, dy, params, seq=ctx.seq, input_grad=want_dx)
        else:
            grads = ctx.module._engine.train_backward(ctx.x, dy, params, seq=ctx.seq, input_grad=want_dx)
        dx = grads.pop("__input__", None)  # local to this rank's shard: never all-reduced
        out = tuple(grads[n] if nd else None for n, nd in zip(ctx.names, need))
        return (None, dx, None) + out


class _EngineFNO(Model):
    """Shared host logic of FNO3d / FNO2d: parameter bookkeeping + engine dispatch."""

    _ndim = 3

    def _make_engine(self, impl="auto"):
        modes = (self.modes1, self.modes2, self.modes3) if self._ndim == 3 else (self.modes1, self.modes2)
        self._engine = FNOEngine(self._ndim, modes, self.n_layers, self.width, self.shape_in, self.shape_out,
                                 padding=self.padding, bn_eps=self.bns[0].eps, impl=impl)

    @property
    def engine(self) -> FNOEngine:
        return self._engine

    def set_impl(self, impl: str):
        """'auto' | 'simt' | 'tc' — which layer kernels the engine uses."""
        compute = getattr(self, "_compute", "f32")
        self._make_engine(impl)
        self._engine.set_compute(compute)

    def set_compute(self, compute: str):
        """'f32' (default) | 'bf16'.  'bf16' = what the reference computes under ``torch.autocast(dtype=bfloat16)``
        (SURVEY F7): Linear / Conv operands in bf16 with fp32 accumulation, everything spectral in fp32 - evaluation
        forward, rollout and the training step (forward + backward GEMMs).  Inside a ``torch.autocast("cuda", dtype=torch.bfloat16)`` context the eval-mode
        forward selects it by itself (and returns a bf16 tensor, as the reference module does)."""
        self._compute = compute
        self._engine.set_compute(compute)

    def _resolve_compute(self):
        """(compute, autocast_active): an active CUDA bf16 autocast context overrides the explicit setting."""
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return "bf16", True
        return getattr(self, "_compute", "f32"), False

    def _engine_state(self):
        sd = {k: v for k, v in self.named_parameters()}
        for i, bn in enumerate(self.bns):
            sd[f"bns.{i}.running_mean"], sd[f"bns.{i}.running_var"] = bn.running_mean, bn.running_var
        key = tuple((t.data_ptr(), t._version) for t in sd.values()) + (getattr(self, "_stats_epoch", 0),)
        return sd, key

    def _check_eval(self):
        if self.training:
            raise RuntimeError("b200fno: rollout() is the evaluation loop of eval.py; call model.eval() first")

    def forward(self, x):
        sd, key = self._engine_state()
        compute, autocast = self._resolve_compute()
        if not self.training:
            self._engine.set_compute(compute)
            y = self._engine.forward(x.float() if autocast else x, sd, key)
            return y.to(torch.bfloat16) if autocast else y  # fc2 is a Linear: bf16 output under autocast
        # train mode: the same switch (bf16: every Linear / Conv GEMM of forward and backward on bf16-rounded operands);
        # the output stays fp32 (autocast would hand the loss a bf16 tensor and cast it straight back to fp32)
        self._engine.set_compute(compute)
        if autocast:
            x = x.float()
        # train mode (train.py:325-329): batch-statistics BatchNorm; differentiable w.r.t. every parameter
        names = [k for k, _ in self.named_parameters()]
        return _TrainForward.apply(self, x, names, *[sd[k] for k in names])

    def rollout(self, x0, affine_a, affine_b, n_steps, out=None, graph=False):
        """Fused eval.py:313-321 loop; see realpdebench_b200.rollout for the full per-batch protocol.
        ``graph=True`` replays a CUDA graph captured once per (batch, n_steps) (FNOEngine.rollout)."""
        self._check_eval()
        sd, key = self._engine_state()
        self._engine.set_compute(self._resolve_compute()[0])
        return self._engine.rollout(x0, affine_a, affine_b, n_steps, sd, key, out=out, graph=graph)

    def train_loss(self, input, target):
        pred = self(input)  # fno.py:131-133
        return mse_loss(pred, target)


class FNO3d(_EngineFNO):
    """fno.py:66-143 — same constructor: FNO3d(modes1, modes2, modes3, n_layers, width, shape_in, shape_out)."""

    _ndim = 3

    def __init__(self, modes1, modes2, modes3, n_layers, width, shape_in, shape_out):
        super().__init__()
        self.modes1, self.modes2, self.modes3 = modes1, modes2, modes3
        self.width = width
        self.shape_in, self.shape_out = tuple(shape_in), tuple(shape_out)
        self.dim_in = shape_in[-1]
        self.dim_out = shape_out[-1] * shape_out[0] // shape_in[0]  # C_out * T_out / T_in (fno.py:86)
        self.padding = 6  # fno.py:87
        self.fc0 = nn.Linear(self.dim_in + 3, self.width)
        self.n_layers = n_layers
        self.spectral_convs, self.convs, self.bns = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(n_layers):  # construction (= RNG) order of fno.py:96-100
            self.spectral_convs.append(SpectralConv3d(width, width, modes1, modes2, modes3))
            self.convs.append(nn.Conv3d(width, width, 1))
            self.bns.append(nn.BatchNorm3d(width))
        self.fc1 = nn.Linear(width, 128)
        self.fc2 = nn.Linear(128, self.dim_out)
        self._make_engine()


class FNO2d(_EngineFNO):
    """FNO-2D of SURVEY.md 8(c): FNO2d(modes1, modes2, n_layers, width, shape_in, shape_out).

    Frames are folded into channels (lift feature t*C_in + c, then grid h, w;
    projection feature t_out*C_out + c); FFT over (H, W) with two weight corners."""

    _ndim = 2

    def __init__(self, modes1, modes2, n_layers, width, shape_in, shape_out):
        super().__init__()
        self.modes1, self.modes2 = modes1, modes2
        self.width = width
        self.shape_in, self.shape_out = tuple(shape_in), tuple(shape_out)
        self.dim_in = shape_in[0] * shape_in[-1]
        self.dim_out = shape_out[0] * shape_out[-1]
        self.padding = 6
        self.fc0 = nn.Linear(self.dim_in + 2, self.width)
        self.n_layers = n_layers
        self.spectral_convs, self.convs, self.bns = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(n_layers):
            self.spectral_convs.append(SpectralConv2d(width, width, modes1, modes2))
            self.convs.append(nn.Conv2d(width, width, 1))
            self.bns.append(nn.BatchNorm2d(width))
        self.fc1 = nn.Linear(width, 128)
        self.fc2 = nn.Linear(128, self.dim_out)
        self._make_engine()
