"""Generate tests/golden/*.npz by executing the REAL reference (build container only).

TEST INFRASTRUCTURE.  Run as `python -m oracle.make_golden` from the repo root in the build
container, where /root/reference exists.  The reference's `src/utils/idealscore.py` is imported
unmodified (behind a matplotlib stub, oracle/ref_loader.py) and its LS / ELS / bbELS modules and
`ScheduledScoreMachine` are executed on CPU (true fp32) on small seeded banks.  Inputs AND outputs
are stored, so the fixtures are self-contained: the GPU box has no /root/reference.

Every case is a dict of arrays:
    kind, bank [N,C,H,W] f32, labels [N] i64, x [1,C,H,W] f32, t, k, label (-1 = None),
    batch_size, max_samples (-1 = None), score [1,C,H,W] f32   (module cases)
    kind, bank, labels, x, scales, label, batch_size, out [1,C,H,W]   (machine cases)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _module(ref, kind, ds, k, bs, max_samples):
    if kind == "ELS":
        return ref.LocalEquivScoreModule(ds, kernel_size=k, batch_size=bs, max_samples=max_samples,
                                         schedule=ref.cosine_noise_schedule)
    if kind == "bbELS":
        return ref.LocalEquivBordersScoreModule(ds, kernel_size=k, batch_size=bs, max_samples=max_samples,
                                                schedule=ref.cosine_noise_schedule)
    if kind == "LS":
        return ref.LocalScoreModule(ds, kernel_size=k, batch_size=bs, max_samples=max_samples,
                                    schedule=ref.cosine_noise_schedule)
    if kind == "IS":
        return ref.IdealScoreModule(ds, batch_size=bs, max_samples=max_samples, schedule=ref.cosine_noise_schedule)
    raise ValueError(kind)


def _one_module_case(ref, kind, c, h, n, k, t, label, bs, ms, seed):
    bank, labels = synthetic_bank(n, c, h, nlabels=4, seed=seed)
    g = torch.Generator().manual_seed(100 + seed)
    j = int(torch.randint(0, n, (1,), generator=g))
    tt = torch.tensor([t])
    beta = float(ref.cosine_noise_schedule(tt))
    x = (1 - beta) ** 0.5 * bank[j:j + 1] + beta ** 0.5 * torch.randn(1, c, h, h, generator=g)
    ds = ref_loader.TensorBank(bank, labels)
    torch.manual_seed(0)
    mod = _module(ref, kind, ds, k, bs, ms)
    lab = None if label is None else torch.tensor([label])
    with torch.no_grad():
        s = mod(tt, x.clone(), label=lab, device=torch.device("cpu"))
    print(f"module {kind:5s} C={c} H={h} N={n} k={k} t={t} label={label}: |score|max={s.abs().max():.3f}")
    return dict(kind=kind, bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(), t=np.float64(t),
                k=np.int64(k), label=np.int64(-1 if label is None else label),
                batch_size=np.int64(bs), max_samples=np.int64(-1 if ms is None else ms), score=s.numpy())


def module_cases(ref):
    cases = []
    # (kind, C, H, N, k, t, label, batch_size, max_samples, bank_seed)
    spec = [
        ("ELS", 3, 12, 40, 3, 0.30, None, 16, None, 0),     # short last batch -> mean quirk
        ("ELS", 3, 12, 40, 5, 0.55, 3, 16, None, 0),        # label filter -> unequal n_b
        ("ELS", 1, 12, 48, 7, 0.80, None, 16, None, 1),
        ("ELS", 1, 16, 32, 3, 0.05, None, 32, None, 2),     # low noise, sharp softmax
        ("ELS", 3, 16, 48, 9, 0.95, 1, 12, 30, 3),          # max_samples break (pre-filter count)
        ("ELS", 3, 8, 24, 5, 0.50, None, 8, None, 4),
        ("bbELS", 3, 12, 40, 3, 0.30, None, 16, None, 0),
        ("bbELS", 3, 12, 40, 5, 0.55, 3, 16, None, 0),
        ("bbELS", 1, 12, 48, 7, 0.80, None, 16, None, 1),
        ("bbELS", 1, 16, 32, 3, 0.05, None, 32, None, 2),
        ("bbELS", 3, 16, 48, 9, 0.95, 1, 12, 30, 3),        # q += batch_size bookkeeping
        ("bbELS", 3, 32, 8, 17, 0.90, None, 8, None, 5),    # CIFAR shape, k=17 (cfg-4 geometry)
        ("LS", 3, 12, 40, 3, 0.30, None, 40, None, 0),      # drivers use one giant batch
        ("LS", 1, 12, 48, 5, 0.55, None, 48, None, 1),
        ("LS", 3, 16, 48, 7, 0.80, 2, 48, None, 3),
        ("LS", 1, 28, 24, 5, 0.40, None, 24, None, 6),      # MNIST native shape (cfg-1 geometry)
        ("LS", 3, 12, 40, 3, 0.30, None, 8, None, 0),       # equal batches: order independent
        ("IS", 3, 12, 40, 3, 0.30, None, 16, None, 0),      # whole-image ideal score (k is ignored)
        ("IS", 1, 16, 48, 3, 0.60, 1, 12, 30, 3),           # label filter + max_samples (post-filter count)
        ("IS", 3, 32, 24, 3, 0.85, None, 24, None, 5),
    ]
    for kind, c, h, n, k, t, label, bs, ms, seed in spec:
        cases.append(_one_module_case(ref, kind, c, h, n, k, t, label, bs, ms, seed))
    # bbELS with k >= H delegates to its internal LS (idealscore.py:163-164); single batch so the
    # hard-coded shuffle=True cannot change the result
    bank, labels = synthetic_bank(16, 3, 8, nlabels=4, seed=7)
    g = torch.Generator().manual_seed(107)
    x = torch.randn(1, 3, 8, 8, generator=g)
    tt = torch.tensor([0.6])
    mod = _module(ref, "bbELS", ref_loader.TensorBank(bank, labels), 9, 16, None)
    with torch.no_grad():
        s = mod(tt, x.clone(), device=torch.device("cpu"), k=9)
    cases.append(dict(kind="bbELS", bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(), t=np.float64(0.6),
                      k=np.int64(9), label=np.int64(-1), batch_size=np.int64(16), max_samples=np.int64(-1),
                      score=s.numpy()))
    # added later (appended so that the earlier files keep their numbers): kernel sizes whose trailing rows go through
    # the mixed K layout of the tensor-core kernel, and IS on an image larger than 32 pixels
    more = [
        ("ELS", 3, 32, 24, 9, 0.60, None, 16, None, 8),
        ("ELS", 3, 32, 16, 13, 0.75, 2, 8, None, 9),
        ("ELS", 1, 28, 24, 11, 0.65, None, 24, None, 10),
        ("ELS", 3, 32, 12, 17, 0.90, None, 12, None, 11),
        ("IS", 3, 40, 12, 3, 0.50, None, 12, None, 12),
    ]
    for kind, c, h, n, k, t, label, bs, ms, seed in more:
        cases.append(_one_module_case(ref, kind, c, h, n, k, t, label, bs, ms, seed))
    # round 2 (appended): bbELS at the CIFAR shape for small and medium kernel sizes, and the kernel sizes of the UNet scales
    # files (checkpoints/scales_*UNet*.pt reach 25 and 27) for ELS and bbELS
    round2 = [
        ("bbELS", 3, 32, 8, 3, 0.15, None, 8, None, 13),
        ("bbELS", 3, 32, 8, 7, 0.40, 1, 4, None, 14),
        ("bbELS", 3, 32, 8, 11, 0.65, None, 8, None, 15),
        ("ELS", 3, 32, 6, 19, 0.92, None, 6, None, 16),
        ("ELS", 3, 32, 6, 25, 0.95, None, 6, None, 17),
        ("ELS", 1, 32, 8, 27, 0.95, None, 8, None, 18),
        ("bbELS", 3, 32, 6, 19, 0.92, None, 6, None, 16),
        ("bbELS", 3, 32, 6, 25, 0.95, None, 6, None, 17),
        ("bbELS", 1, 32, 8, 27, 0.95, None, 8, None, 18),
    ]
    for kind, c, h, n, k, t, label, bs, ms, seed in round2:
        cases.append(_one_module_case(ref, kind, c, h, n, k, t, label, bs, ms, seed))
    return cases


def machinex_cases(ref):
    """Trajectories outside the plain (kind, scales) pattern: an IdealScoreModule backbone without a scales list
    (ScheduledScoreMachine then runs default_time_steps steps and passes k=None, idealscore.py:85-95), and LS with
    batch_size < N plus a label, where the reference opens a freshly shuffled DataLoader at every evaluation (two draws
    from the global torch RNG each, idealscore.py:489,521) -- the test replays it from the same torch seed."""
    cases = []
    # IS, no scales
    bank, labels = synthetic_bank(24, 3, 12, nlabels=3, seed=21)
    mod = _module(ref, "IS", ref_loader.TensorBank(bank, labels), 3, 8, None)
    machine = ref.ScheduledScoreMachine(mod, in_channels=3, imsize=12, default_time_steps=6, score_backbone=True)
    x = torch.randn(1, 3, 12, 12, generator=torch.Generator().manual_seed(221))
    with torch.no_grad():
        out = machine(x.clone(), label=torch.tensor([1]), device=torch.device("cpu"))
    cases.append(dict(kind="IS", bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(), scales=np.zeros(0, np.int64),
                      nsteps=np.int64(6), label=np.int64(1), batch_size=np.int64(8), torch_seed=np.int64(-1),
                      out=out.numpy()))
    print(f"machinex IS no scales: |out|max={out.abs().max():.3f}")
    # LS, shuffled batches of unequal post-filter size
    bank, labels = synthetic_bank(40, 3, 12, nlabels=3, seed=22)
    scales = [3, 3, 5, 5, 7, 7]
    mod = _module(ref, "LS", ref_loader.TensorBank(bank, labels), 3, 16, None)
    machine = ref.ScheduledScoreMachine(mod, in_channels=3, imsize=12, scales=scales, score_backbone=True)
    x = torch.randn(1, 3, 12, 12, generator=torch.Generator().manual_seed(222))
    torch.manual_seed(1234)
    with torch.no_grad():
        out = machine(x.clone(), label=torch.tensor([2]), device=torch.device("cpu"))
    cases.append(dict(kind="LS", bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(),
                      scales=np.asarray(scales, dtype=np.int64), nsteps=np.int64(len(scales)), label=np.int64(2),
                      batch_size=np.int64(16), torch_seed=np.int64(1234), out=out.numpy()))
    print(f"machinex LS shuffled: |out|max={out.abs().max():.3f}")
    return cases


def machine_cases(ref):
    cases = []
    spec = [
        ("ELS", 3, 12, 32, [3, 3, 3, 5, 5, 7], None, 16, 10),
        ("ELS", 1, 12, 36, [3, 3, 5, 5, 7, 7, 9, 9], 2, 12, 11),
        ("bbELS", 3, 12, 32, [3, 3, 3, 5, 5, 7], None, 16, 10),
        ("bbELS", 1, 12, 36, [3, 3, 5, 5, 7, 7, 9, 13], 2, 12, 11),   # last k >= H -> LS fallback, 1 batch/class
        ("LS", 3, 12, 32, [3, 3, 3, 5, 5, 7], None, 32, 10),
        # added later (appended): the headline geometry with the kernel sizes of the shipped CIFAR schedule, and a
        # MNIST-shape run with a non-monotone schedule like scales_MNIST_ResNet_circular
        ("ELS", 3, 32, 20, [3, 3, 5, 7, 9, 11, 13, 15, 17, 17], 1, 8, 12),
        ("ELS", 1, 28, 24, [3, 3, 5, 7, 11, 11, 9, 7, 3], None, 10, 13),
    ]
    for kind, c, h, n, scales, label, bs, seed in spec:
        bank, labels = synthetic_bank(n, c, h, nlabels=3, seed=seed)
        ds = ref_loader.TensorBank(bank, labels)
        if kind == "bbELS" and max(scales) >= h:
            bs = n  # keep the internal shuffled LS single-batch so the golden is order independent
        torch.manual_seed(0)
        mod = _module(ref, kind, ds, 3, bs, None)
        machine = ref.ScheduledScoreMachine(mod, in_channels=c, imsize=h, scales=scales, score_backbone=True)
        g = torch.Generator().manual_seed(200 + seed)
        x = torch.randn(1, c, h, h, generator=g)
        lab = None if label is None else torch.tensor([label])
        with torch.no_grad():
            out = machine(x.clone(), label=lab, device=torch.device("cpu"))
        cases.append(dict(kind=kind, bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(),
                          scales=np.asarray(scales, dtype=np.int64), label=np.int64(-1 if label is None else label),
                          batch_size=np.int64(bs), out=out.numpy()))
        print(f"machine {kind:5s} C={c} H={h} N={n} scales={scales} label={label}: |out|max={out.abs().max():.3f}")
    return cases


def ddpm_cases(ref):
    """The stochastic branch of the reference sampler, `DDIM.sample(ddpm=True)` (src/models.py:48-64), driven by the
    reference's own analytic score module: the backbone handed to DDIM returns eps = -sqrt(beta) * score.  The Gaussian
    noise comes from torch's global CPU generator (randn_like, one draw per step); the test regenerates the same draws
    from `torch_seed` and injects them."""
    models = ref_loader.load_models()
    cases = []
    for kind, c, h, n, k, nsteps, label, bs, seed in [("ELS", 3, 12, 32, 5, 8, None, 16, 31), ("bbELS", 1, 16, 24, 3, 6, 1, 8, 32)]:
        bank, labels = synthetic_bank(n, c, h, nlabels=3, seed=seed)
        mod = _module(ref, kind, ref_loader.TensorBank(bank, labels), k, bs, None)

        class EpsBackbone(torch.nn.Module):
            def forward(self, t, x, label=None):
                s = mod(t.cpu(), x, label=label, device=torch.device("cpu"))
                return -ref.cosine_noise_schedule(t.cpu())[:, None, None, None] ** 0.5 * s

        sampler = models.DDIM(backbone=EpsBackbone(), in_channels=c, noise_schedule=ref.cosine_noise_schedule, default_imsize=h)
        x = torch.randn(1, c, h, h, generator=torch.Generator().manual_seed(300 + seed))
        lab = None if label is None else torch.tensor([label])
        torch.manual_seed(4000 + seed)
        with torch.no_grad():
            out = sampler.sample(batch_size=1, x=x.clone(), nsteps=nsteps, label=lab, device=torch.device("cpu"), ddpm=True)
        cases.append(dict(kind=kind, bank=bank.numpy(), labels=labels.numpy(), x=x.numpy(), nsteps=np.int64(nsteps),
                          k=np.int64(k), label=np.int64(-1 if label is None else label), batch_size=np.int64(bs),
                          torch_seed=np.int64(4000 + seed), out=out.numpy()))
        print(f"ddpm {kind} C={c} H={h} N={n} k={k} nsteps={nsteps}: |out|max={out.abs().max():.3f}")
    return cases


def grad_cases(ref):
    """Vector-Jacobian products of the score obtained by autograd THROUGH the reference modules (what the reference's
    exterior-derivative callers do, src/utils/exterior_derivative.py:68-79): grad = d/dx sum(score(x) * g)."""
    cases = []
    for kind, c, h, n, k, t, label, bs, seed in [("ELS", 3, 8, 12, 3, 0.4, None, 8, 41), ("ELS", 1, 12, 10, 5, 0.7, 1, 4, 42),
                                                 ("bbELS", 3, 10, 10, 3, 0.5, None, 10, 43), ("bbELS", 1, 12, 12, 5, 0.3, None, 6, 44),
                                                 ("LS", 3, 10, 16, 3, 0.6, None, 16, 45), ("IS", 1, 8, 12, 3, 0.5, None, 12, 46)]:
        bank, labels = synthetic_bank(n, c, h, nlabels=3, seed=seed)
        mod = _module(ref, kind, ref_loader.TensorBank(bank, labels), k, bs, None)
        gen = torch.Generator().manual_seed(500 + seed)
        x = torch.randn(1, c, h, h, generator=gen).requires_grad_(True)
        g = torch.randn(1, c, h, h, generator=gen)
        lab = None if label is None else torch.tensor([label])
        s = mod(torch.tensor([t]), x, label=lab, device=torch.device("cpu"))
        (grad,) = torch.autograd.grad((s * g).sum(), x)
        cases.append(dict(kind=kind, bank=bank.numpy(), labels=labels.numpy(), x=x.detach().numpy(), g=g.numpy(), t=np.float64(t),
                          k=np.int64(k), label=np.int64(-1 if label is None else label), batch_size=np.int64(bs),
                          score=s.detach().numpy(), grad=grad.numpy()))
        print(f"grad {kind} C={c} H={h} N={n} k={k}: |grad|max={grad.abs().max():.3f}")
    return cases


def schedule_cases(ref):
    t = torch.arange(0, 21, dtype=torch.float32) / 20
    return dict(t=t.numpy(), cosine=ref.cosine_noise_schedule(t).numpy(),
                exponential=ref.exponential_schedule(t).numpy())


def scales_files():
    import glob
    out = {}
    for f in sorted(glob.glob(os.path.join(ref_loader.REF_DIR, "checkpoints", "scales_*.pt"))):
        out[os.path.basename(f)[:-3]] = np.asarray(torch.load(f, weights_only=False), dtype=np.int64)
    return out


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    force = "--force" in sys.argv          # default: keep the committed fixtures, write only the ones that are missing

    def save(name, arrays):
        path = os.path.join(OUT, name)
        if force or not os.path.exists(path):
            np.savez_compressed(path, **arrays)
            print("wrote", name)

    for i, c in enumerate(module_cases(ref)):
        save(f"module_{i:02d}_{c['kind']}.npz", c)
    for i, c in enumerate(machine_cases(ref)):
        save(f"machine_{i:02d}_{c['kind']}.npz", c)
    for i, c in enumerate(machinex_cases(ref)):
        save(f"machinex_{i:02d}_{c['kind']}.npz", c)
    for i, c in enumerate(grad_cases(ref)):
        save(f"grad_{i:02d}_{c['kind']}.npz", c)
    for i, c in enumerate(ddpm_cases(ref)):
        save(f"ddpm_{i:02d}_{c['kind']}.npz", c)
    save("schedule.npz", schedule_cases(ref))
    save("scales.npz", scales_files())


if __name__ == "__main__":
    main()
