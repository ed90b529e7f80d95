"""GPU measurement (not a pytest file): ELS on 64x64x3 images (BASELINE config 5 geometry), per-evaluation device time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalEquivScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    dev = torch.device("cuda")
    bank, labels = synthetic_bank(n, 3, 64, nlabels=1, seed=0)
    mod = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule, precision="auto")
    eng = mod.engine(dev)
    sel = mod.selection(None)
    for B in (1, 4):
        x = noisy_query(bank, 0.5, B, seed=1).to(dev)
        for k, t in ((3, 0.1), (7, 0.4), (11, 0.7), (17, 0.9)):
            beta_val = float(cosine_noise_schedule(torch.tensor([t])))
            beta = torch.full((B,), beta_val, device=dev)
            passes = eng.passes_for(k, beta_val)
            eng.umma_partials("circular", x, beta, k, sel, passes)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                eng.umma_partials("circular", x, beta, k, sel, passes)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 2
            pairs = B * 64 * 64 * n * (64 - k + 1) ** 2
            print(f"ELS 64x64x3 N={n} B={B} k={k} passes={passes}: {ms:.2f} ms  {pairs / ms * 1e3:.3e} pairs/s  "
                  f"{pairs * 2 * k * k * 3 / ms * 1e-9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
