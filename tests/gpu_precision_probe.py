"""GPU diagnostic: how far does a single bf16 pass move the denoised estimate at every step of the headline
workload (full 50k CIFAR-shape bank, class conditional)?  Evidence for the `precision="auto"` rule.
    python tests/gpu_precision_probe.py > gpurun_out/precision_probe.log"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalEquivScoreModule, ScheduledScoreMachine, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.scales import load_scales  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402


def main():
    bank, labels = synthetic_bank(50000, 3, 32, seed=0)
    scales = load_scales("CIFAR10_ResNet_zeros_conditional")
    mod = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule,
                                precision="f16x2")
    machine = ScheduledScoreMachine(mod, in_channels=3, imsize=32, scales=scales)
    eng = mod.engine("cuda")
    for label in (0, 5):
        x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(label)).cuda()
        out, rec = machine.trajectory(x, label=label, device="cuda")
        sel = mod.selection(label)
        print(f"label {label}: n_sel={sel[2]}")
        for r in rec:
            beta = torch.full((2,), r["beta"], device="cuda")
            mu1 = torch.empty_like(x)
            P = eng.combine(eng.umma_partials("circular", r["x"], beta, r["k"], sel, 1, tag="probe"))
            eng.finalize(P, r["x"], beta, mu1, None)
            d = (mu1 - r["mu"]).abs()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.umma_partials("circular", r["x"], beta, r["k"], sel, 2, tag="probe")
            torch.cuda.synchronize()
            t2 = time.perf_counter() - t0
            t0 = time.perf_counter()
            eng.umma_partials("circular", r["x"], beta, r["k"], sel, 1, tag="probe")
            torch.cuda.synchronize()
            t1 = time.perf_counter() - t0
            a_over_b = (1 - r["beta"]) ** 0.5 / r["beta"]
            print(f"  step i={r['i']:2d} k={r['k']:2d} beta={r['beta']:.4f} a/beta={a_over_b:8.3f} "
                  f"|mu1-mu2| max={d.max():.3e} rms={d.pow(2).mean().sqrt():.3e}   t(2 pass)={t2*1e3:.2f} ms t(1 pass)={t1*1e3:.2f} ms",
                  flush=True)


if __name__ == "__main__":
    main()
