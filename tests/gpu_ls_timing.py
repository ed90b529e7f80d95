"""GPU diagnostic: device time of the LS kernels alone (engine-level launches, no host work in between)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402


def main():
    dev = torch.device("cuda")
    for (C, h, n) in ((1, 28, 60000), (1, 32, 60000), (3, 32, 50000)):
        bank, labels = synthetic_bank(n, C, h, seed=0)
        mod = LocalScoreModule((bank, labels), kernel_size=5, batch_size=n, schedule=cosine_noise_schedule)
        eng = mod.engine(dev)
        sel = mod.selection(None)
        for B in (1, 4, 16):
            x = noisy_query(bank, 0.3, B, seed=1).to(dev)
            beta = torch.full((B,), 0.3, device=dev)
            mu = torch.empty_like(x)
            for k in (5, 17):
                def step():
                    P = eng.ls_partials(x, beta, k, sel)
                    eng.finalize(eng.combine(P), x, beta, mu, None)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                a.record()
                for _ in range(reps):
                    step()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                a.record()
                for _ in range(reps):
                    eng.ls_partials(x, beta, k, sel)
                b.record()
                torch.cuda.synchronize()
                msk = a.elapsed_time(b) / reps
                gb = n * C * h * h * 4 / 1e9
                print(f"LS C={C} H={h} N={n} k={k} B={B}: eval {ms*1e3:.1f} us (partials kernel {msk*1e3:.1f} us) "
                      f"bank {gb*1e3:.0f} MB -> {gb/msk*1e3:.0f} GB/s of 6553 measured peak, "
                      f"{B*h*h*n/ms*1e3:.3e} pairs/s", flush=True)


if __name__ == "__main__":
    main()
