"""Sample generation driver with the reference's flags and on-disk layout.

Counterpart of `/root/reference/scripts/els_script.py` (which does not parse as shipped: SyntaxError at :124): builds one
analytic score module + `ScheduledScoreMachine`, then writes per-sample tensors

    results/<expname>/seeds/NNNN.pt          [1,C,H,W] fp32   the initial noise
    results/<expname>/<idealname>/NNNN.pt    [1,C,H,W] fp32   the machine's output          (els_script.py:200-204)
    results/<expname>/labels/NNNN.pt         int64 (1,)       only with --conditional

which is what `scripts/eval_script.py:40-64` reads.  Re-running resumes at the first missing index; `--fill` recomputes
outputs for existing seeds (els_script.py:145-166); `--force_overwrite` starts over (:181-189).

Datasets are not downloadable here: `--dataset` names a synthetic stand-in of the right shape (synthetic.SHAPES) or
`--bankfile` points at a `torch.save((images [N,C,H,W] in [-1,1], labels [N]))` file.

    python -m convolutional_diffusion_b200.els_script --dataset cifar10 --scoremoduletype ELS --conditional --numiters 8
"""
from __future__ import annotations

import argparse
import os
import shutil

import torch

from . import (IdealScoreModule, LocalEquivBordersScoreModule, LocalEquivScoreModule, LocalScoreModule,
               ScheduledScoreMachine, cosine_noise_schedule)
from .scales import SCALES, load_scales
from .synthetic import SHAPES, synthetic_dataset


def build_module(kind, dataset, n, image_size, channels, scorebatchsize, max_samples, shuffle, **kw):
    """Same constructor choices as els_script.py:69-96 (LS and IS use one batch covering the dataset)."""
    schedule = cosine_noise_schedule
    if kind == "ELS":
        return LocalEquivScoreModule(dataset, batch_size=scorebatchsize, image_size=image_size, channels=channels,
                                     schedule=schedule, shuffle=shuffle, max_samples=max_samples, **kw)
    if kind == "bbELS":
        return LocalEquivBordersScoreModule(dataset, batch_size=scorebatchsize, image_size=image_size,
                                            channels=channels, schedule=schedule, max_samples=max_samples, **kw)
    if kind == "LS":
        return LocalScoreModule(dataset, image_size=image_size, batch_size=n, show_plots=False, schedule=schedule, **kw)
    if kind == "IS":
        return IdealScoreModule(dataset, image_size=image_size, batch_size=n, schedule=schedule, **kw)
    raise ValueError(f"Unknown scoremoduletype: {kind}")


def default_scales(dataset_name, conditional):
    """Mirror of the auto-detection order of els_script.py:99-117 over the shipped scales lists."""
    up = {"mnist": "MNIST", "mnist28": "MNIST", "cifar10": "CIFAR10", "fashionmnist": "FashionMNIST",
          "celeba": "CelebA", "celeba64": "CelebA"}.get(dataset_name, dataset_name)
    for cand in (f"{up}_ResNet_zeros_conditional", f"{up}_ResNet_zeros_conditonal", f"{up}_ResNet_zeros",
                 f"{up}_UNet_zeros_conditional", f"{up}_UNet_zeros_conditonal", f"{up}_UNet_zeros"):
        if cand in SCALES:
            return cand
    raise FileNotFoundError("No scales found. Please specify --scalesfile")


def first_missing(seedpath, spath, numiters):
    for i in range(numiters):
        if not (os.path.exists(os.path.join(seedpath, f"{i:04d}.pt")) and os.path.exists(os.path.join(spath, f"{i:04d}.pt"))):
            return i
    return numiters


def main(argv=None):
    ap = argparse.ArgumentParser(description="Generate_Data")
    ap.add_argument("--expname", type=str, default=None)
    ap.add_argument("--idealname", type=str, default="els_outputs")
    ap.add_argument("--dataset", type=str, default="mnist", help=f"synthetic stand-in, one of {sorted(SHAPES)}")
    ap.add_argument("--bankfile", type=str, default=None, help="torch.save((images, labels)) file instead of --dataset")
    ap.add_argument("--banksize", type=int, default=None, help="number of synthetic images (default: the dataset's size)")
    ap.add_argument("--scoremoduletype", type=str, default="bbELS")
    ap.add_argument("--conditional", action="store_true", default=False)
    ap.add_argument("--scalesfile", type=str, default=None)
    ap.add_argument("--scorebatchsize", type=int, default=256)
    ap.add_argument("--fill", action="store_true", default=False)
    ap.add_argument("--numiters", type=int, default=100)
    ap.add_argument("--nsteps", type=int, default=20)
    ap.add_argument("--nlabels", type=int, default=10)
    ap.add_argument("--samplebatch", type=int, default=1,
                    help="samples generated per machine call (extension; labels are drawn per sample)")
    ap.add_argument("--force_overwrite", action="store_true", default=False)
    ap.add_argument("--max_samples", type=int, default=100000)
    ap.add_argument("--shuffle", action="store_true", default=False)
    ap.add_argument("--results", type=str, default="./results")
    ap.add_argument("--precision", type=str, default="auto")
    args = ap.parse_args(argv)

    if args.bankfile:
        images, labels = torch.load(args.bankfile, weights_only=False)
        name = os.path.splitext(os.path.basename(args.bankfile))[0]
        channels, image_size = int(images.shape[1]), int(images.shape[-1])
    else:
        images, labels, meta = synthetic_dataset(args.dataset, n=args.banksize)
        name, channels, image_size = meta["name"], meta["num_channels"], meta["image_size"]
    dataset = (images, labels)
    expname = args.expname or f"dataset_{name}_option_{args.scoremoduletype}" + ("_conditional" if args.conditional else "")

    mod = build_module(args.scoremoduletype, dataset, len(images), image_size, channels, args.scorebatchsize,
                       args.max_samples, args.shuffle, precision=args.precision)
    scales = load_scales(args.scalesfile if args.scalesfile else default_scales(name, args.conditional))
    machine = ScheduledScoreMachine(mod, in_channels=channels, imsize=image_size, noise_schedule=cosine_noise_schedule,
                                    score_backbone=True, scales=scales)
    device = torch.device("cuda")

    dpath = os.path.join(args.results, expname)
    seedpath, spath, lpath = os.path.join(dpath, "seeds"), os.path.join(dpath, args.idealname), os.path.join(dpath, "labels")

    if args.fill:
        if not os.path.isdir(dpath) or not os.path.isdir(seedpath):
            raise FileNotFoundError(f"Required directories not found: {dpath} or {seedpath}")
        os.makedirs(spath, exist_ok=True)
        i = 0
        while os.path.exists(os.path.join(seedpath, f"{i:04d}.pt")):
            seed = torch.load(os.path.join(seedpath, f"{i:04d}.pt"), weights_only=False)
            label = torch.load(os.path.join(lpath, f"{i:04d}.pt"), weights_only=False) if args.conditional else None
            if not os.path.exists(os.path.join(spath, f"{i:04d}.pt")):
                out = machine(seed.clone(), label=label, device=device)
                torch.save(out.cpu(), os.path.join(spath, f"{i:04d}.pt"))
            i += 1
        return i

    if os.path.isdir(dpath) and not args.force_overwrite:
        start = first_missing(seedpath, spath, args.numiters)
    else:
        if os.path.isdir(dpath):
            shutil.rmtree(dpath)
        start = 0
    for pth in (seedpath, spath) + ((lpath,) if args.conditional else ()):
        os.makedirs(pth, exist_ok=True)
    i = start
    while i < args.numiters:
        nb = max(1, min(args.samplebatch, args.numiters - i))     # samples per machine call (the reference: 1)
        seeds = torch.randn(nb, channels, image_size, image_size, device=device)
        labels = torch.randint(0, args.nlabels, (nb,)) if args.conditional else None
        outs = machine(seeds.clone(), label=labels, device=device)
        for j in range(nb):                                       # same per-sample files as the reference writes
            torch.save(seeds[j:j + 1].cpu(), os.path.join(seedpath, f"{i + j:04d}.pt"))
            torch.save(outs[j:j + 1].cpu(), os.path.join(spath, f"{i + j:04d}.pt"))
            if args.conditional:
                torch.save(labels[j:j + 1].clone(), os.path.join(lpath, f"{i + j:04d}.pt"))
        i += nb
    return args.numiters


if __name__ == "__main__":
    main()
