"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE.

Build-container only: imports the unmodified reference package from
``/root/reference`` (behind empty ``matplotlib`` / ``h5py`` stub modules, which
the hot path never calls) and records small input/output vectors.  The GPU box
has no ``/root/reference``; tests there read the committed ``*.pt`` files.

    python tests/golden/make_golden.py

What is recorded (torch 2.11.0 CPU):
  kat_a.pt       FNO3d forward, recipe of SURVEY.md section 4 (KAT-A)
  kat_b.pt       SpectralConv3d forward (KAT-B)
  fno3d_odd.pt   FNO3d with odd grid sizes, T_out = 2*T_in (r=2 unfold) and train-mode BN
  rollout.pt     eval.py:297-326 executed verbatim (source lines exec'd) on the
                 reference FNO3d + GaussianNormalizer: plain (C_in==C_out),
                 controlled (C_in = C_out+2) and RangeNormalizer cases
  spectral2d.pt  MWT sparseKernelFT2d spectral math (2-D semantics pin)
  train3d.pt     train.py:321-334 executed on the reference FNO3d in .train() mode (loss, gradients after the
                 first backward, losses and state_dict after 3 Adam + StepLR steps); a plain case and one with
                 overlapping spectral corners + T_out = 2*T_in      (python tests/golden/make_golden.py train)
  metrics.pt     realpdebench/utils/metrics.py eval_metrics run on random fields: whole batch, chunked (batch_size=2,
                 c=2 of 3 channels), single channel, and a real-data-like case with an all-zero pressure channel
                                                                    (python tests/golden/make_golden.py metrics)
  surrogate.pt   data/generate_surrogate_data.py:63-86 (the per-file loop body, source lines exec'd verbatim) on a small
                 reference FNO3d + GaussianNormalizer and a seeded 9-frame 128 x 128 x 15 trajectory (the script
                 hard-codes that frame shape); the trajectory is regenerated from its seed, only the output is stored
  siblings.pt    MWT sparseKernelFT3d / sparseKernelFT2d and the Galerkin SpectralConv3d forwards (modules imported)
                                                                    (python tests/golden/make_golden.py widening)
"""
import os
import sys
import textwrap
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from realpdebench.model.fno import FNO3d, SpectralConv3d
    from realpdebench.data.data_normalizer import GaussianNormalizer, RangeNormalizer
    from realpdebench.utils.metrics import mse_loss
    return FNO3d, SpectralConv3d, GaussianNormalizer, RangeNormalizer, mse_loss


def randomize_bn(m, seed=123):
    g = torch.Generator().manual_seed(seed)
    for bn in m.bns:
        c = bn.weight.numel()
        bn.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
        bn.weight.data.copy_(torch.rand(c, generator=g) + 0.5)
        bn.bias.data.copy_(torch.randn(c, generator=g) * 0.1)


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def eval_loop_source():
    """The reference's rollout body, eval.py:297-326, as an exec-able string."""
    with open(os.path.join(REF, "realpdebench", "eval.py")) as f:
        lines = f.readlines()
    body = "".join(lines[296:326])  # 1-based 297..326
    assert "b = input.size(0)" in lines[296] and "postprocess(input, target)" in lines[325]
    return textwrap.dedent(body)


def run_reference_rollout(model, normalizer, input, target, n_auto, mse_loss):
    ns = dict(torch=torch, model=model, data_normalizer=normalizer, input=input, target=target,
              args=types.SimpleNamespace(N_autoregressive=n_auto), mse_loss=mse_loss,
              normalized_test_loss=0.0)
    with torch.no_grad():
        exec(eval_loop_source(), ns)
    return ns["pred"], ns["target"], ns["normalized_test_loss"], ns["preds"]


def make_train():
    """The reference training step, train.py:321-334 (the script body, restated here line by line around the
    unmodified reference module + torch.optim.Adam / StepLR exactly as train.py:290-292 builds them)."""
    torch.set_num_threads(1)
    FNO3d, _, _, _, _ = import_reference()
    cases = {}
    for name, ctor, bsz, clip in (
            ("plain", (2, 3, 4, 2, 8, (4, 8, 12, 3), (4, 8, 12, 3)), 2, 0.0),
            ("overlap_r2", (5, 3, 2, 2, 6, (2, 9, 7, 2), (4, 9, 7, 2)), 3, 0.5)):
        torch.manual_seed(40)
        m = FNO3d(*ctor)
        randomize_bn(m, 41)
        sd0 = sd_of(m)
        torch.manual_seed(42)
        batches = [(torch.randn(bsz, *ctor[5]), torch.randn(bsz, *ctor[6])) for _ in range(3)]
        optimizer = torch.optim.Adam(m.parameters(), lr=1e-3)  # train.py:290
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=2, gamma=0.5)  # train.py:292
        losses, grads0, pred0 = [], None, None
        for it, (input, target) in enumerate(batches):
            m.train()  # train.py:322
            optimizer.zero_grad()  # :325
            if it == 0:
                with torch.no_grad():
                    sd_keep = sd_of(m)
                    pred0 = m(input)  # train-mode forward (updates running stats) ...
                    m.load_state_dict(sd_keep)  # ... undone, so the recorded step starts from sd0
            loss = m.train_loss(input, target).mean()  # :328
            loss.backward()  # :329
            if it == 0:
                grads0 = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
            if clip > 0:
                torch.nn.utils.clip_grad_norm_(m.parameters(), clip)  # :330-331
            optimizer.step()  # :333
            scheduler.step()  # :334
            losses.append(loss.item())
        print("train", name, losses)
        cases[name] = dict(ctor=ctor, sd0=sd0, batches=batches, lr=1e-3, step_size=2, clip=clip, losses=losses,
                           grads0=grads0, pred0=pred0, sd_final=sd_of(m))
    torch.save(cases, os.path.join(HERE, "train3d.pt"))
    print("train3d.pt", os.path.getsize(os.path.join(HERE, "train3d.pt")))


def make_metrics():
    """eval_metrics (utils/metrics.py:24-131) of the reference itself on small random fields."""
    torch.set_num_threads(1)
    import_reference()
    from realpdebench.utils.metrics import eval_metrics
    cases = []
    torch.manual_seed(60)
    for shape, c, bs in (((3, 8, 10, 12, 3), 3, None), ((5, 12, 8, 14, 3), 2, 2), ((2, 6, 6, 6, 1), 1, None),
                         ((4, 10, 16, 12, 3), 2, None)):
        pred, target = torch.randn(*shape), torch.randn(*shape)
        target = target + 0.5 * torch.sin(torch.arange(shape[1]).float()).reshape(1, -1, 1, 1, 1)
        pred = target + 0.3 * pred
        out = eval_metrics(pred, target, c, bs)
        cases.append(dict(pred=pred, target=target, c=c, batch_size=bs, out=torch.stack([torch.as_tensor(o).float() for o in out])))
        print("metrics", shape, c, bs, [round(float(o), 5) for o in out])
    torch.save(cases, os.path.join(HERE, "metrics.pt"))


def surrogate_inputs(seed=70, n=9):
    """Seeded stand-in for ``hf['measured_data']`` (generate_surrogate_data.py:48): [n, 128, 128, 15] float32."""
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, n).reshape(n, 1, 1, 1)
    return (torch.randn(n, 128, 128, 15, generator=g) * 0.5 + torch.sin(6.0 * t)).numpy()


def make_widening():
    """the surrogate materialisation loop and the sibling spectral layers, run from the reference."""
    import importlib.util
    import numpy as np
    torch.set_num_threads(1)
    FNO3d, _, GaussianNormalizer, _, _ = import_reference()

    # ---- generate_surrogate_data.py:63-86, exec'd on a small model ---------
    with open(os.path.join(REF, "realpdebench", "data", "generate_surrogate_data.py")) as f:
        lines = f.readlines()
    assert "for i in range(0, traj_numerical.shape[0]-1, batch_size*step):" in lines[62]
    assert "pred_list.append(pred_traj_batch.reshape(-1, 128, 128)[[-1]].cpu().numpy())" in lines[85]
    body = textwrap.dedent("".join(lines[62:86]))
    step, batch_size, sub_s = 2, 2, 1
    ctor = (2, 3, 3, 2, 8, (step, 128, 128, 17), (step, 128, 128, 1))
    torch.manual_seed(71)
    model = FNO3d(*ctor).eval()
    randomize_bn(model, 72)
    g = torch.Generator().manual_seed(73)
    norm = object.__new__(GaussianNormalizer)
    norm.device = "cpu"
    norm.mean_inputs, norm.std_inputs = torch.randn(17, generator=g) * 0.1, torch.rand(17, generator=g) + 0.5
    norm.mean_targets, norm.std_targets = torch.randn(1, generator=g) * 0.1, torch.rand(1, generator=g) + 0.5
    traj = surrogate_inputs()
    ns = dict(torch=torch, np=np, model=model, data_normalizer=norm, traj_numerical=traj, step=step,
              batch_size=batch_size, sub_s=sub_s, gas_ratio=40, equivalence_ratio=0.85, device="cpu", pred_list=[])
    exec(body, ns)
    pred_traj = np.concatenate(ns["pred_list"], axis=0)  # :88
    print("surrogate", pred_traj.shape, float(pred_traj.sum()), float(np.abs(pred_traj).sum()))
    torch.save(dict(ctor=ctor, sd=sd_of(model), seed=70, n=9, step=step, batch_size=batch_size, sub_s=sub_s,
                    gas_ratio=40, equivalence_ratio=0.85,
                    norm=dict(mean_inputs=norm.mean_inputs, std_inputs=norm.std_inputs,
                              mean_targets=norm.mean_targets, std_targets=norm.std_targets),
                    traj_checksum=float(np.abs(traj).sum()), pred_traj=torch.from_numpy(pred_traj)),
               os.path.join(HERE, "surrogate.pt"))

    # ---- sibling spectral layers -------------------------------------------
    from realpdebench.model.MWT_libs.models import sparseKernelFT2d, sparseKernelFT3d
    spec = importlib.util.spec_from_file_location(  # the package __init__ needs IPython; the layer file does not
        "_galerkin_layers", os.path.join(REF, "realpdebench", "model", "galerkin_transformer_libs", "layers.py"))
    gl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gl)
    cases = {}
    torch.manual_seed(80)
    for name, k, alpha, c, shape in (("ft3d", 2, 3, 1, (2, 8, 6, 10)),      # all modes kept
                                     ("ft3d_clipped", 2, 4, 2, (2, 4, 5, 8))):  # l1 = 3 < modes, overlapping corners
        m = sparseKernelFT3d(k, alpha, c).eval()
        x = torch.randn(*shape, c, k * k)
        with torch.no_grad():
            y = m(x)
        cases[name] = dict(kind="ft3d", k=k, alpha=alpha, c=c, x=x, y=y, sd=sd_of(m))
        print("siblings", name, tuple(y.shape), y.sum().item())
    for name, k, alpha, c, shape in (("ft2d", 3, 5, 1, (2, 16, 20)), ("ft2d_clipped", 2, 6, 1, (3, 8, 10))):
        m = sparseKernelFT2d(k, alpha, c).eval()
        x = torch.randn(*shape, c, k * k)
        with torch.no_grad():
            y = m(x)
        cases[name] = dict(kind="ft2d", k=k, alpha=alpha, c=c, x=x, y=y, sd=sd_of(m))
        print("siblings", name, tuple(y.shape), y.sum().item())
    m = gl.SpectralConv3d(5, 6, 3, 2, 4).eval()  # (in_dim, out_dim, modes_x, modes_y, modes_t)
    x = torch.randn(2, 5, 7, 9, 8)
    with torch.no_grad():
        y = m(x)
    cases["galerkin3d"] = dict(kind="galerkin3d", ctor=(5, 6, 3, 2, 4), x=x, y=y, sd=sd_of(m))
    print("siblings galerkin3d", tuple(y.shape), y.sum().item())
    torch.save(cases, os.path.join(HERE, "siblings.pt"))
    for f in ("surrogate.pt", "siblings.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


def main():
    torch.set_num_threads(1)
    FNO3d, SpectralConv3d, GaussianNormalizer, RangeNormalizer, mse_loss = import_reference()

    # ---- KAT-A -----------------------------------------------------------
    torch.manual_seed(0)
    m = FNO3d(2, 4, 4, 2, 8, (10, 16, 32, 3), (10, 16, 32, 3)).eval()
    randomize_bn(m)
    torch.manual_seed(1)
    x = torch.randn(2, 10, 16, 32, 3)
    with torch.no_grad():
        y = m(x)
    print("KAT-A", y.sum().item(), y.abs().sum().item(), y[0, 0, 0, 0].tolist())
    torch.save(dict(sd=sd_of(m), x=x, y=y, ctor=(2, 4, 4, 2, 8, (10, 16, 32, 3), (10, 16, 32, 3))),
               os.path.join(HERE, "kat_a.pt"))

    # ---- KAT-B -----------------------------------------------------------
    torch.manual_seed(2)
    s = SpectralConv3d(4, 5, 2, 3, 4)
    torch.manual_seed(3)
    z = torch.randn(2, 4, 9, 10, 12)
    with torch.no_grad():
        o = s(z)
    print("KAT-B", o.sum().item(), o.abs().sum().item(), o[0, 0, 0, 0, :3].tolist())
    torch.save(dict(w=[getattr(s, f"weights{k}").detach().clone() for k in (1, 2, 3, 4)], z=z, o=o),
               os.path.join(HERE, "kat_b.pt"))

    # ---- odd sizes, r = 2, train-mode BN -----------------------------------
    torch.manual_seed(10)
    m = FNO3d(3, 3, 2, 3, 6, (5, 9, 7, 2), (10, 9, 7, 2))
    randomize_bn(m, 77)
    torch.manual_seed(11)
    x = torch.randn(3, 5, 9, 7, 2)
    sd0 = sd_of(m)
    with torch.no_grad():
        y_eval = m.eval()(x)
        y_train = m.train()(x)
    torch.save(dict(sd=sd0, sd_after_train=sd_of(m), x=x, y_eval=y_eval, y_train=y_train,
                    ctor=(3, 3, 2, 3, 6, (5, 9, 7, 2), (10, 9, 7, 2))),
               os.path.join(HERE, "fno3d_odd.pt"))

    # ---- rollout, reference eval.py lines executed verbatim -------------------
    cases = {}
    for name, c_in, c_out, kind in (("plain", 3, 3, "gaussian"), ("controlled", 5, 3, "gaussian"),
                                    ("range", 3, 3, "range")):
        torch.manual_seed(20)
        m = FNO3d(2, 3, 4, 2, 8, (4, 8, 12, c_in), (4, 8, 12, c_out)).eval()
        randomize_bn(m, 5)
        g = torch.Generator().manual_seed(4321)
        if kind == "gaussian":
            n = object.__new__(GaussianNormalizer)
            n.device = "cpu"
            n.mean_inputs, n.std_inputs = torch.randn(c_in, generator=g) * 0.1, torch.rand(c_in, generator=g) + 0.5
            n.mean_targets, n.std_targets = torch.randn(c_out, generator=g) * 0.1, torch.rand(c_out, generator=g) + 0.5
            stats = dict(mean_inputs=n.mean_inputs, std_inputs=n.std_inputs,
                         mean_targets=n.mean_targets, std_targets=n.std_targets)
        else:
            n = object.__new__(RangeNormalizer)
            n.device = "cpu"
            n.max_inputs, n.max_targets = torch.rand(c_in, generator=g) + 1.0, torch.rand(c_out, generator=g) + 1.0
            stats = dict(max_inputs=n.max_inputs, max_targets=n.max_targets)
        torch.manual_seed(21)
        n_auto = 3
        inp = torch.randn(2, 4, 8, 12, c_in)
        tgt = torch.randn(2, 4 * n_auto, 8, 12, c_out)
        tgt[..., -1] = 0  # real-data convention: unmeasured pressure channel is all zero (fluid_dataset.py:357-359)
        pred, tgt_dn, loss, preds = run_reference_rollout(m, n, inp, tgt, n_auto, mse_loss)
        print("rollout", name, pred.shape, loss)
        cases[name] = dict(sd=sd_of(m), ctor=(2, 3, 4, 2, 8, (4, 8, 12, c_in), (4, 8, 12, c_out)), kind=kind,
                           stats=stats, input=inp, target=tgt, n_auto=n_auto, pred=pred, target_dn=tgt_dn,
                           loss=loss, states=[p.clone() for p in preds])
    torch.save(cases, os.path.join(HERE, "rollout.pt"))

    # ---- 2-D spectral semantics: MWT sparseKernelFT2d -------------------------
    from realpdebench.model.MWT_libs.models import sparseKernelFT2d
    torch.manual_seed(30)
    k, c, alpha = 2, 2, 5  # channels = c*k^2 = 8, modes = 5
    sk = sparseKernelFT2d(k, alpha, c)
    ch = c * k * k
    with torch.no_grad():
        sk.Lo.weight.copy_(torch.eye(ch))
        sk.Lo.bias.zero_()
        xx = torch.randn(2, 14, 18, c, k * k)
        # forward ends with relu -> Lo(identity); the spectral map S is linear, so
        # relu(S x) - relu(S(-x)) = S x recovers it from the unmodified module.
        yy = sk(xx) - sk(-xx)
    torch.save(dict(w1=sk.weights1.detach().clone(), w2=sk.weights2.detach().clone(),
                    x=xx.view(2, 14, 18, ch).permute(0, 3, 1, 2).contiguous(),
                    y=yy.view(2, 14, 18, ch).permute(0, 3, 1, 2).contiguous()),
               os.path.join(HERE, "spectral2d.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        make_train()
    elif len(sys.argv) > 1 and sys.argv[1] == "metrics":
        make_metrics()
    elif len(sys.argv) > 1 and sys.argv[1] == "widening":
        make_widening()
    else:
        main()
        make_train()
        make_metrics()
        make_widening()
