"""Debug helper (not a pytest file): module goldens through the FMA and P.V epilogues, error maps against the oracle."""
import glob
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from oracle import score_oracle as so  # noqa: E402


def main():
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "module_*_ELS.npz")))
    for f in files:
        c = dict(np.load(f))
        label = None if int(c["label"]) < 0 else int(c["label"])
        ms = None if int(c["max_samples"]) < 0 else int(c["max_samples"])
        k = int(c["k"])
        beta = float(so.cosine_beta(float(c["t"])))
        a = (1 - beta) ** 0.5
        idx, logw = so.select_bank("ELS", c["labels"], label, int(c["batch_size"]), ms)
        _, mu_o = so.score("ELS", c["x"][0], c["bank"][idx], beta, k, logw)
        out = {}
        for variant in ("fma", "pv"):
            ds = (torch.from_numpy(c["bank"]).float(), torch.from_numpy(c["labels"]).long())
            mod = cd.LocalEquivScoreModule(ds, kernel_size=k, batch_size=int(c["batch_size"]), max_samples=ms,
                                           schedule=cd.cosine_noise_schedule)
            mod.engine("cuda").els_variant = variant
            x = torch.from_numpy(c["x"]).cuda()
            s = mod(torch.tensor([float(c["t"])]), x, label=None if label is None else torch.tensor([label]), device="cuda")
            mu = (c["x"][0].astype(np.float64) + beta * s.cpu().double().numpy()[0]) / a
            out[variant] = mu
            err = np.abs(mu - mu_o)
            print(f"{os.path.basename(f)} k={k} beta={beta:.3f} n_sel={len(idx)} {variant}: max err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
                  f"  #>1e-4: {(err > 1e-4).sum()} of {err.size}")
        d = np.abs(out["pv"] - out["fma"])
        print("   pv-fma max", d.max(), "rows with diff>1e-4:", sorted(set(np.nonzero(d > 1e-4)[1].tolist())),
              "cols:", sorted(set(np.nonzero(d > 1e-4)[2].tolist())))


if __name__ == "__main__":
    main()
