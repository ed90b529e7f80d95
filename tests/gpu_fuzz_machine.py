"""Fuzz of ScheduledScoreMachine (CUDA-graph trajectory) against the float64 oracle: random schedules of kernel sizes,
image sizes 8..32, ELS / bbELS / LS, labels, ragged DataLoader batches (not a pytest file).
Usage: python tests/gpu_fuzz_machine.py [seed] [trials]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402
from oracle import score_oracle as so  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rng = np.random.default_rng(seed)
    worst_db = 1e9
    for trial in range(trials):
        kind = ["ELS", "bbELS", "LS"][trial % 3]
        C = int(rng.choice([1, 3]))
        H = int(rng.integers(8, 33))
        nsteps = int(rng.integers(5, 13))
        # kernel sizes grow with the noise level like the shipped schedules; scales[0] is never read
        kmax = min(H - (1 if kind == "bbELS" else 0), 17)
        ks = sorted(int(rng.choice([v for v in range(3, kmax + 1, 2)])) for _ in range(nsteps))
        N = int(rng.integers(6, 30))
        bs = N if kind == "LS" else int(rng.integers(3, N + 1))
        label = None if rng.random() < 0.5 else int(rng.integers(0, 2))
        bank, labels = synthetic_bank(N, C, H, nlabels=2, seed=900 + trial)
        if label is not None and not bool((labels == label).any()):
            label = int(labels[0])
        cls = {"ELS": cd.LocalEquivScoreModule, "bbELS": cd.LocalEquivBordersScoreModule, "LS": cd.LocalScoreModule}[kind]
        mod = cls((bank, labels), kernel_size=3, batch_size=bs, schedule=cd.cosine_noise_schedule)
        machine = cd.ScheduledScoreMachine(mod, in_channels=C, imsize=H, scales=ks)
        x = torch.randn(2, C, H, H, generator=torch.Generator().manual_seed(trial))
        lab = None if label is None else torch.tensor([label])
        out = machine(x.cuda(), label=lab, device=torch.device("cuda")).cpu().double().numpy()
        out2 = machine(x.cuda(), label=lab, device=torch.device("cuda")).cpu().double().numpy()     # graph replay
        assert np.array_equal(out, out2), "graph replay differs"
        idx, logw = so.select_bank(kind, labels.numpy(), label, bs, None)
        psnr = 1e9
        for b in range(2):
            ref = so.run_machine(kind, x[b].numpy(), bank.numpy()[idx], ks, logw)
            psnr = min(psnr, 10 * np.log10(4.0 / max(float(np.mean((out[b] - ref) ** 2)), 1e-30)))
        worst_db = min(worst_db, psnr)
        print(f"trial {trial}: {kind} C={C} H={H} N={N} bs={bs} label={label} scales={ks} PSNR={psnr:.1f} dB", flush=True)
        assert psnr >= 50.0
    print(f"machine fuzz worst PSNR {worst_db:.1f} dB")


if __name__ == "__main__":
    main()
