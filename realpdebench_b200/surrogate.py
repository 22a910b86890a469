"""Surrogate materialisation on the engine.

``materialize_surrogate`` is the per-trajectory body of the reference script
``realpdebench/data/generate_surrogate_data.py:58-88``: a trained FNO3d maps windows of ``step`` simulated frames
(+ two constant parameter channels) to the measured observable; windows are independent (no autoregression), the
last frame comes from one extra window over the final ``step`` frames.

Engine form: the windows of one chunk are one ``model.rollout(x, a, b, 1)`` call, so the target de-normalisation
(``postprocess``, data_normalizer.py:57-62) is the projection kernel's per-channel affine and never a separate pass;
the host->device copy of chunk i+1 and the device->host copy of prediction i-1 run on side streams under the forward
of chunk i (pinned staging buffers).  The parameter-channel concat and the input normalisation are torch elementwise
kernels on the device (plumbing).  No CPU path: the model must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch


def postprocess_affine(normalizer, c_out: int, device):
    """Per-channel (a, b) with ``postprocess(., p)[1] == p * a + b`` for the reference normalisers."""
    if hasattr(normalizer, "mean_targets"):  # GaussianNormalizer, data_normalizer.py:57-62
        a, b = normalizer.std_targets[..., :c_out], normalizer.mean_targets[..., :c_out]
    elif hasattr(normalizer, "max_targets"):  # RangeNormalizer, data_normalizer.py:125-130
        a = normalizer.max_targets[..., :c_out]
        b = torch.zeros_like(a)
    else:  # IdentityNormalizer, data_normalizer.py:11-17
        a, b = torch.ones(c_out), torch.zeros(c_out)
    return (a.to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
            b.to(device=device, dtype=torch.float32).reshape(-1).contiguous())


def window_plan(n_frames: int, step: int, batch_size: int):
    """Chunk boundaries of generate_surrogate_data.py:63 plus the tail window (:77).

    Returns ``([(start, stop), ...], (tail_start, tail_stop), n_pred)``; ``n_pred`` is the length of ``pred_traj``
    (:88) for a single-channel target.  A chunk whose length is not a multiple of ``step`` cannot be windowed: the
    reference fails in ``reshape`` (:65) and so does this function."""
    if step < 1 or batch_size < 1:
        raise ValueError("step and batch_size must be >= 1")
    if n_frames < step:
        raise RuntimeError(f"trajectory has {n_frames} frames, fewer than one window of {step}")
    chunks = []
    for i in range(0, n_frames - 1, batch_size * step):
        j = min(i + batch_size * step, n_frames)
        if (j - i) % step:
            raise RuntimeError(f"shape '[-1, {step}, ...]' is invalid for a chunk of {j - i} frames "
                               f"(frames {i}:{j} of {n_frames}; generate_surrogate_data.py:65)")
        chunks.append((i, j))
    return chunks, (n_frames - step, n_frames), sum(j - i for i, j in chunks) + 1


def materialize_surrogate(model, data_normalizer, traj_numerical, gas_ratio, equivalence_ratio, step: int = 10,
                          batch_size: int = 50, sub_s: int = 1, _pipes=None) -> np.ndarray:
    """``pred_traj`` of generate_surrogate_data.py:88 for one trajectory.

    * ``model``: engine ``FNO3d`` in eval mode on a CUDA device, ``shape_in = (step, H, W, C + 2)``,
      ``shape_out = (step, H, W, 1)`` (the script's model, :27-35)
    * ``data_normalizer``: the reference ``GaussianNormalizer`` / ``RangeNormalizer`` / ``IdentityNormalizer`` (or any
      object with the same statistics attributes)
    * ``traj_numerical``: host array ``[n, H0, W0, C]`` (``hf['measured_data']``, :48), subsampled by ``sub_s`` (:65)
    """
    traj = np.asarray(traj_numerical)
    if traj.ndim != 4:
        raise ValueError(f"traj_numerical must be [n, H, W, C], got shape {traj.shape}")
    if traj.dtype != np.float32:
        traj = traj.astype(np.float32)  # torch.tensor(..., dtype=torch.float) of :65
    traj = traj[:, ::sub_s, ::sub_s]
    n, h, w, c = traj.shape
    chunks, tail, n_pred = window_plan(n, step, batch_size)
    device = next(model.parameters()).device
    if _pipes is None:
        if device.type != "cuda":
            raise RuntimeError("b200fno: materialize_surrogate runs on the CUDA engine only (no CPU fallback); "
                               "move the model to a CUDA device")
        _pipes = _CudaPipes(device)
    c_in, c_out = model.shape_in[-1], model.shape_out[-1]
    if tuple(model.shape_in) != (step, h, w, c + 2):
        raise RuntimeError(f"b200fno: model.shape_in {tuple(model.shape_in)} does not match windows "
                           f"{(step, h, w, c + 2)} (measured channels + gas ratio + equivalence ratio)")
    a, b = postprocess_affine(data_normalizer, c_out, device)
    mean_in = std_in = max_in = None
    if hasattr(data_normalizer, "mean_inputs"):
        mean_in = data_normalizer.mean_inputs[..., :c_in].to(device=device, dtype=torch.float32)
        std_in = data_normalizer.std_inputs[..., :c_in].to(device=device, dtype=torch.float32)
    elif hasattr(data_normalizer, "max_inputs"):
        max_in = data_normalizer.max_inputs[..., :c_in].to(device=device, dtype=torch.float32)

    windows = chunks + [tail]
    max_frames = max(j - i for i, j in windows)
    host_in = [_pipes.pinned(max_frames * h * w * c) for _ in range(2)]
    host_out = _pipes.pinned((n_pred + step - 1) * h * w * c_out)
    h2d_done = [None, None]  # per pinned staging buffer: its last host->device copy

    def stage(k):
        i, j = windows[k]
        if h2d_done[k & 1] is not None:
            _pipes.host_wait(h2d_done[k & 1])  # the copy of window k-2 has left this buffer (it preceded forward k-2)
        hb = host_in[k & 1][:(j - i) * h * w * c]
        hb.view(j - i, h, w, c).numpy()[...] = traj[i:j]  # strided (sub_s) gather into pinned memory
        d, ev = _pipes.h2d(hb)
        h2d_done[k & 1] = ev
        return d, ev

    out_pos, pieces = 0, []
    nxt = stage(0)
    with torch.no_grad():
        for k, (i, j) in enumerate(windows):
            d, ev = nxt
            _pipes.compute_wait(ev, d)
            nb = (j - i) // step
            x = torch.empty((nb, step, h, w, c_in), dtype=torch.float32, device=device)
            x[..., :c] = d.view(nb, step, h, w, c)
            x[..., c] = gas_ratio  # :66-68
            x[..., c + 1] = equivalence_ratio
            if mean_in is not None:  # preprocess, data_normalizer.py:50-55
                x = (x - mean_in) / std_in
            elif max_in is not None:  # :118-123
                x = x / max_in
            p = model.rollout(x, a, b, 1)  # forward + postprocess (:71-72) in one engine call
            numel = p.numel()
            _pipes.d2h(host_out[out_pos:out_pos + numel], p)  # under the next chunk's forward
            pieces.append((out_pos, numel))
            out_pos += numel
            if k + 1 < len(windows):
                nxt = stage(k + 1)  # host gather + H2D of the next chunk while this chunk's forward runs
    _pipes.finish()
    flat = host_out.numpy()
    parts = [flat[o:o + m].reshape(-1, h, w) for o, m in pieces]  # :74
    parts[-1] = parts[-1][[-1]]  # :86, only the last frame of the tail window
    return np.concatenate(parts, axis=0)  # :88


class _CudaPipes:
    """Copy streams and pinned staging for :func:`materialize_surrogate` (the only device plumbing it needs)."""

    def __init__(self, device):
        self.device = device
        self.copy_in, self.copy_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self.cur = torch.cuda.current_stream(device)

    def pinned(self, numel: int) -> torch.Tensor:
        return torch.empty(numel, dtype=torch.float32).pin_memory()

    def h2d(self, host: torch.Tensor):
        with torch.cuda.stream(self.copy_in):
            d = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_in)
        return d, ev

    def host_wait(self, ev) -> None:
        ev.synchronize()

    def compute_wait(self, ev, d: torch.Tensor) -> None:
        self.cur.wait_event(ev)
        d.record_stream(self.cur)

    def d2h(self, host: torch.Tensor, p: torch.Tensor) -> None:
        ready = torch.cuda.Event()
        ready.record(self.cur)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ready)
            host.copy_(p.reshape(-1), non_blocking=True)
        p.record_stream(self.copy_out)

    def finish(self) -> None:
        self.copy_out.synchronize()
