"""One batch of the reference evaluation loop, eval.py:296-326, on the engine.

``rollout`` takes the same objects the reference script has in scope (model,
data_normalizer, input, target, N_autoregressive) and returns what the script
keeps: the de-normalised prediction and target (eval.py:325-326, 342-343) and
the normalised loss term (eval.py:323).  The N-step loop with its
postprocess / cat / preprocess glue (eval.py:313-319) is one engine call.
"""
from __future__ import annotations

import torch


def rollout_affine(normalizer, c_in: int, c_out: int, device):
    """Per-channel (a, b) with  preprocess(postprocess(p)) == p*a + b  (eval.py:316,318):
    the model output is de-normalised with TARGET statistics and re-normalised with INPUT statistics."""
    if hasattr(normalizer, "mean_inputs"):  # GaussianNormalizer, data_normalizer.py:50-62
        mi, si = normalizer.mean_inputs[..., :c_out], normalizer.std_inputs[..., :c_out]
        mt, st = normalizer.mean_targets[..., :c_out], normalizer.std_targets[..., :c_out]
        a, b = st / si, (mt - mi) / si
    elif hasattr(normalizer, "max_inputs"):  # RangeNormalizer, data_normalizer.py:118-130
        a = normalizer.max_targets[..., :c_out] / normalizer.max_inputs[..., :c_out]
        b = torch.zeros_like(a)
    else:  # IdentityNormalizer
        a, b = torch.ones(c_out), torch.zeros(c_out)
    return (a.to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
            b.to(device=device, dtype=torch.float32).reshape(-1).contiguous())


def _rollout_device(model, data_normalizer, input, target, N_autoregressive: int, c: int, denorm_target: bool = True):
    """eval.py:311-326 on the device without any host synchronisation: returns the de-normalised prediction, the
    de-normalised target and the normalised loss as a 0-d device tensor."""
    b = input.size(0)
    c_in, c_out = input.shape[-1], target.shape[-1]
    with torch.no_grad():
        input, target = data_normalizer.preprocess(input, target)  # eval.py:311 (H2D + affine)
        a, bb = rollout_affine(data_normalizer, c_in, c_out, input.device)
        pred = model.rollout(input, a, bb, N_autoregressive)  # eval.py:313-322, parameter channels already dropped
        loss = torch.nn.functional.mse_loss(pred[..., :c], target[..., :c], reduction='none') \
            .reshape(b, -1).mean()  # eval.py:323
        _, pred = data_normalizer.postprocess(input, pred)  # eval.py:325
        if denorm_target:
            _, target = data_normalizer.postprocess(input, target)  # eval.py:326
    return pred, target, loss


def _count_unmeasured(target) -> int:
    """eval.py:298-302: target channels that are identically zero (unmeasured in the real-world data)."""
    return sum(int(torch.all(target[..., c_] == 0)) for c_ in range(target.shape[-1]))


def rollout(model, data_normalizer, input, target, N_autoregressive: int, unmeasured_c=None):
    """Returns ``(pred, target, normalized_loss, unmeasured_c)`` for one batch.

    * ``pred``, ``target``: de-normalised tensors on the device (eval.py:325-326)
    * ``normalized_loss``: python float added to ``normalized_test_loss`` (eval.py:323)
    * ``unmeasured_c``: all-zero target channels, computed on the first batch
      like eval.py:298-302 and passed back in for the following ones.
    """
    if unmeasured_c is None:
        unmeasured_c = _count_unmeasured(target)
    c = target.shape[-1] - unmeasured_c
    pred, target, loss = _rollout_device(model, data_normalizer, input, target, N_autoregressive, c)
    return pred, target, loss.item(), unmeasured_c


def rollout_stream(model, data_normalizer, batches, N_autoregressive: int, unmeasured_c=None, to_host: bool = False,
                   host_ring: int = 2):
    """The evaluation loop eval.py:296-343 over an iterable of HOST ``(input, target)`` batches.

    Yields ``(pred, target, normalized_loss)`` per batch like :func:`rollout`, but stages the host->device
    copy of batch i+1 on a side stream while batch i is being rolled out, so PCIe time overlaps compute
    (pinned host tensors make the copies truly asynchronous; pageable ones still work, synchronously).

    ``to_host=True`` is eval.py:342-343 (``pred_list.append(pred.cpu())``) as part of the pipeline: the
    de-normalised prediction of batch i is copied to a pinned host buffer on a third stream while batch i+1 is
    copied in and rolled out (PCIe is full duplex), and the generator yields host tensors ``(pred, target, loss)``.
    ``target`` is then the batch's own host tensor: preprocess followed by postprocess is the identity
    (data_normalizer.py:50-62), so the round trip eval.py:326/343 makes through the device is skipped.  The pinned
    prediction buffers form a ring of ``host_ring`` entries: a yielded ``pred`` is valid until ``host_ring - 1`` further
    items have been requested (copy it, or pass a larger ring, to keep more of them like ``pred_list`` does).
    """
    it = iter(batches)
    try:
        first = next(it)
    except StopIteration:
        return
    device = next(model.parameters()).device
    copy_stream = torch.cuda.Stream(device=device)
    out_stream = torch.cuda.Stream(device=device) if to_host else None
    ring, loss_ring = [], []

    def stage(batch):
        inp, tgt = batch
        with torch.cuda.stream(copy_stream):
            d_in = inp.to(device, non_blocking=True)
            d_tg = tgt.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d_in, d_tg, ev, tgt

    def finish(job):
        pred_h, tgt_h, loss_h, ev = job
        ev.synchronize()
        return pred_h, tgt_h, float(loss_h)

    nxt = stage(first)
    pending = None  # (host pred, host target, host loss, event) of the previous batch, D2H possibly still in flight
    i = 0
    while nxt is not None:
        d_in, d_tg, ev, tgt_host = nxt
        batch = next(it, None)
        nxt = stage(batch) if batch is not None else None  # copy of the next batch overlaps this rollout
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        d_in.record_stream(cur), d_tg.record_stream(cur)
        if not to_host:
            pred, target, loss, unmeasured_c = rollout(model, data_normalizer, d_in, d_tg, N_autoregressive, unmeasured_c)
            yield pred, target, loss
            continue
        if unmeasured_c is None:
            cur.synchronize()
            unmeasured_c = _count_unmeasured(d_tg)
        c = d_tg.shape[-1] - unmeasured_c
        pred, _, loss = _rollout_device(model, data_normalizer, d_in, d_tg, N_autoregressive, c, denorm_target=False)
        slot = i % host_ring
        if len(ring) <= slot:
            ring.append(torch.empty(pred.shape, dtype=pred.dtype, pin_memory=True))
            loss_ring.append(torch.empty((), dtype=torch.float32, pin_memory=True))
        elif ring[slot].shape != pred.shape:  # last, smaller batch of a DataLoader without drop_last
            ring[slot] = torch.empty(pred.shape, dtype=pred.dtype, pin_memory=True)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(done)
            ring[slot].copy_(pred, non_blocking=True)
            loss_ring[slot].copy_(loss, non_blocking=True)
            pred.record_stream(out_stream), loss.record_stream(out_stream)
            fin = torch.cuda.Event()
            fin.record(out_stream)
        job = (ring[slot], tgt_host, loss_ring[slot], fin)
        if pending is not None:  # hand out batch i-1 only now: batch i is already enqueued behind it
            yield finish(pending)
        pending = job
        i += 1
    if pending is not None:
        yield finish(pending)
