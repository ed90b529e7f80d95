"""Profiling target (not a pytest file): one score evaluation per listed kernel size on the headline workload
(50k CIFAR-shape bank, class 0, batch 4).  Usage:  python tests/gpu_profile_target.py 3 17"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalEquivScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402


def main():
    warm = "--warm" in sys.argv          # one untimed launch + the mean of 5 (not for ncu captures)
    ks = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [3, 17]
    bank, labels = synthetic_bank(50000, 3, 32, seed=0)
    mod = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule,
                                precision="auto")
    eng = mod.engine("cuda")
    sel = mod.selection(0)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(0)).cuda()
    tmap = {3: 0.10, 5: 0.25, 7: 0.45, 9: 0.60, 11: 0.70, 13: 0.75, 15: 0.80, 17: 0.90}
    for k in ks:
        t = tmap.get(k, 0.5)
        beta_val = float(cosine_noise_schedule(torch.tensor([t])))
        beta = torch.full((4,), beta_val, device="cuda")
        passes = eng.passes_for(k, beta_val)
        eng.bank.norm_plane(k)
        eng.bank.strip8()
        reps = 1
        if warm:
            eng.umma_partials("circular", x, beta, k, sel, passes)
            reps = 5
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.umma_partials("circular", x, beta, k, sel, passes)
        b.record()
        torch.cuda.synchronize()
        print(f"k={k} passes={passes} n_sel={sel[2]} {a.elapsed_time(b) / reps:.3f} ms")


if __name__ == "__main__":
    main()
