"""Per-step mu error of the headline schedule against the float64 oracle on a small bank (not a pytest file).
Usage: python tests/gpu_step_errors.py [seed ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from convolutional_diffusion_b200.scales import load_scales  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402
from oracle import score_oracle as so  # noqa: E402


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [77]
    scales = load_scales("CIFAR10_ResNet_zeros_conditional")
    for seed in seeds:
        bank, labels = synthetic_bank(60, 3, 32, nlabels=3, seed=21 + seed)
        label, bs = 1, 64
        for precision in ("auto", "f16"):
            mod = cd.LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=bs, schedule=cd.cosine_noise_schedule,
                                           precision=precision)
            machine = cd.ScheduledScoreMachine(mod, in_channels=3, imsize=32, scales=scales)
            x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(seed))
            out, rec = machine.trajectory(x.cuda(), label=torch.tensor([label]), device="cuda")
            idx, logw = so.select_bank("ELS", labels.numpy(), label, bs, None)
            sub = bank.numpy()[idx]
            errs = []
            for r in rec:
                mu_o = so.els_mu(r["x"][0].cpu().numpy(), sub, r["beta"], r["k"], logw)
                errs.append((r["k"], float(np.sqrt(1 - r["beta"]) / r["beta"]),
                             float(np.max(np.abs(r["mu"][0].cpu().double().numpy() - mu_o)))))
            print(f"seed={seed} precision={precision}: " + " ".join(f"k{k}:{g:.2g}:{e:.1e}" for k, g, e in errs))


if __name__ == "__main__":
    main()
