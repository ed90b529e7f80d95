"""Probe one ELS geometry against the float64 oracle in every precision / layout mode (not a pytest file).
Usage: python tests/gpu_case_probe.py C H k N t [t ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402
from oracle import score_oracle as so  # noqa: E402


def main():
    C, H, k, N = (int(a) for a in sys.argv[1:5])
    ts = [float(a) for a in sys.argv[5:]] or [0.5]
    bank, labels = synthetic_bank(N, C, H, nlabels=3, seed=100)
    for t in ts:
        beta = float(so.cosine_beta(t))
        x = noisy_query(bank, beta, 1, seed=0)
        idx, logw = so.select_bank("ELS", labels.numpy(), None, N, None)
        _, mu_o = so.score("ELS", x[0].numpy(), bank.numpy()[idx], beta, k, logw)
        row = []
        for name, kw, env in (("simt", dict(use_tensor_cores=False), {}), ("f16x2", dict(precision="f16x2"), {}),
                              ("f16x2-vert", dict(precision="f16x2"), {"CDS_ELS_MIXED": "0"}),
                              ("f16x2-noAtmem", dict(precision="f16x2"), {"CDS_A_TMEM": "0"}),
                              ("f16", dict(precision="f16"), {}), ("auto", dict(precision="auto"), {})):
            for e, v in env.items():
                os.environ[e] = v
            mod = cd.LocalEquivScoreModule((bank, labels), kernel_size=k, batch_size=N, schedule=cd.cosine_noise_schedule, **kw)
            s = mod(torch.tensor([t]), x.cuda(), device=torch.device("cuda")).cpu().double().numpy()[0]
            for e in env:
                os.environ.pop(e)
            mu = (s * beta + x[0].double().numpy()) / np.sqrt(1 - beta)
            row.append(f"{name}:{np.max(np.abs(mu - mu_o)):.1e}")
        a_b = np.sqrt(1 - beta) / beta
        print(f"C={C} H={H} k={k} N={N} t={t} a/beta={a_b:.3g} (a/beta)k^2C={a_b * k * k * C:.3g}: " + " ".join(row))


if __name__ == "__main__":
    main()
