"""GPU measurement sweep over the BASELINE.json configs (not a pytest file): per-evaluation device times of the
score modules at full bank sizes.     python tests/gpu_config_sweep.py > gpurun_out/config_sweep.log"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import (LocalEquivScoreModule, LocalEquivBordersScoreModule, LocalScoreModule,  # noqa: E402
                                           ScheduledScoreMachine, cosine_noise_schedule)
from convolutional_diffusion_b200.scales import load_scales  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def pairs(kind, h, k, n):
    d = k // 2
    if kind == "LS" or (kind == "bbELS" and k >= h):
        return h * h * n
    if kind == "ELS":
        return h * h * n * (h - k + 1) ** 2
    i = h - 2 * d
    return n * (i ** 4 + 4 * d * i * i + 4 * d * d)


def main():
    dev = torch.device("cuda")
    out = []
    # cfg-1: LS, MNIST shape (28x28x1 and the reference's resized 32x32x1), 60k bank, k=5, B=1 and B=10
    for h in (28, 32):
        bank, labels = synthetic_bank(60000, 1, h, seed=0)
        mod = LocalScoreModule((bank, labels), kernel_size=5, batch_size=60000, schedule=cosine_noise_schedule)
        for B in (1, 10):
            x = noisy_query(bank, 0.3, B, seed=1).to(dev)
            ms = timed(lambda: mod(torch.full((B,), 0.4), x, device=dev))
            p = B * pairs("LS", h, 5, 60000)
            gbs = 60000 * h * h * 4 / ms * 1e-6
            out.append(dict(cfg="cfg1 LS", H=h, C=1, N=60000, k=5, B=B, ms=ms, pairs_per_s=p / ms * 1e3, bank_GBps=gbs))
            print(out[-1], flush=True)
        # cfg-2: ELS MNIST shape, zero padding and circular, kernel sizes 3..15
        els = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, channels=1,
                                    schedule=cosine_noise_schedule, query_pad="zeros", precision="auto")
        for k, t in ((3, 0.1), (7, 0.35), (11, 0.6), (15, 0.9)):
            x = noisy_query(bank, 0.3, 1, seed=2).to(dev)
            ms = timed(lambda: els(torch.tensor([t]), x, device=dev, k=k), reps=2)
            p = pairs("ELS", h, k, 60000)
            out.append(dict(cfg="cfg2 ELS zeros", H=h, C=1, N=60000, k=k, B=1, ms=ms, pairs_per_s=p / ms * 1e3,
                            tflops=p * 2 * k * k / ms * 1e-9))
            print(out[-1], flush=True)
        if h == 32:
            machine = ScheduledScoreMachine(els, in_channels=1, imsize=32, scales=load_scales("MNIST_ResNet_zeros"))
            xs = torch.randn(1, 1, 32, 32, device=dev)
            ms = timed(lambda: machine(xs, device=dev), reps=1)
            out.append(dict(cfg="cfg2 ELS zeros trajectory (19 evals, unconditional 60k)", ms=ms))
            print(out[-1], flush=True)
        del mod, els, bank
    # cfg-4: bbELS vs ELS, CIFAR shape, k=17, 50k bank unconditional
    bank, labels = synthetic_bank(50000, 3, 32, seed=0)
    x = noisy_query(bank, 0.9, 1, seed=3).to(dev)
    for name, cls in (("bbELS", LocalEquivBordersScoreModule), ("ELS", LocalEquivScoreModule)):
        mod = cls((bank, labels), kernel_size=17, batch_size=64, schedule=cosine_noise_schedule, precision="auto")
        ms = timed(lambda: mod(torch.tensor([0.9]), x, device=dev), reps=2)
        p = pairs(name, 32, 17, 50000)
        out.append(dict(cfg="cfg4 " + name, H=32, C=3, N=50000, k=17, B=1, ms=ms, pairs_per_s=p / ms * 1e3,
                        tflops=p * 2 * 867 / ms * 1e-9))
        print(out[-1], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
