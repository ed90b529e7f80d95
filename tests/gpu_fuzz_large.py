"""ELS / bbELS / LS / IS fuzz on larger and odd image sizes (33..64 pixels, band staging, partial 8-column blocks, B > 1 with
per-sample noise levels) against the float64 oracle (not a pytest file; minutes of oracle time).
Usage: python tests/gpu_fuzz_large.py [seed] [trials] [--float-bank] [--small]
--float-bank: banks that are not on the 8-bit grid (two-plane strip8, no mixed K layout); --small: 8..32 pixels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402
from oracle import score_oracle as so  # noqa: E402


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    seed = int(pos[0]) if len(pos) > 0 else 0
    trials = int(pos[1]) if len(pos) > 1 else 20
    float_bank, small = "--float-bank" in sys.argv, "--small" in sys.argv
    rng = np.random.default_rng(seed)
    worst = 0.0
    for trial in range(trials):
        kind = ["bbELS", "ELS", "ELS", "LS", "IS"][trial % 5]
        C = int(rng.choice([1, 2, 3])) if kind == "ELS" else int(rng.choice([1, 3]))
        H = int(rng.integers(8, 33)) if small else int(rng.integers(33, 65))
        k = int(rng.choice([v for v in range(3, 33, 2) if v <= H - (1 if kind == "bbELS" else 0)]))
        N = int(rng.integers(3, 9))
        B = int(rng.integers(1, 4))
        ts = [float(rng.uniform(0.05, 0.98)) for _ in range(B)]
        betas = [float(so.cosine_beta(t)) for t in ts]
        for b in range(B):      # stay in the regime fp32 dot products resolve (see test_random_geometries_against_oracle)
            while np.sqrt(1 - betas[b]) / betas[b] * k * k * C > 2e4:
                ts[b] = min(0.98, ts[b] + 0.05)
                betas[b] = float(so.cosine_beta(ts[b]))
        bank, labels = synthetic_bank(N, C, H, nlabels=2, seed=500 + trial)
        if float_bank:
            bank = (bank + 0.37 * torch.rand(bank.shape, generator=torch.Generator().manual_seed(trial)) / 127.5).clamp(-1, 1)
        g = torch.Generator().manual_seed(trial)
        x = torch.stack([np.sqrt(1 - betas[b]) * bank[int(torch.randint(0, N, (1,), generator=g))]
                         + np.sqrt(betas[b]) * torch.randn(C, H, H, generator=g) for b in range(B)]).float()
        if kind == "IS":
            mod = cd.IdealScoreModule((bank, labels), batch_size=N, schedule=cd.cosine_noise_schedule)
        else:
            cls = {"ELS": cd.LocalEquivScoreModule, "bbELS": cd.LocalEquivBordersScoreModule, "LS": cd.LocalScoreModule}[kind]
            mod = cls((bank, labels), kernel_size=k, batch_size=N, schedule=cd.cosine_noise_schedule)
        t0 = time.time()
        s = mod(torch.tensor(ts), x.cuda(), device=torch.device("cuda")).cpu().double().numpy()
        idx, logw = so.select_bank(kind, labels.numpy(), None, N, None)
        err = 0.0
        for b in range(B):
            _, mu_o = so.score(kind, x[b].numpy(), bank.numpy()[idx], betas[b], k, logw)
            mu = (s[b] * betas[b] + x[b].double().numpy()) / np.sqrt(1 - betas[b])
            err = max(err, float(np.max(np.abs(mu - mu_o))))
        worst = max(worst, err)
        print(f"trial {trial}: {kind} C={C} H={H} k={k} N={N} B={B} t={[round(t, 2) for t in ts]} err={err:.2e} ({time.time() - t0:.1f}s)",
              flush=True)
        assert err < 1e-3
    print(f"large fuzz worst mu error {worst:.2e}")


if __name__ == "__main__":
    main()
