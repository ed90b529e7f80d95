"""Debug / A-B script (GPU box): LS on the tensor cores (csrc/ls_umma.cu) vs the SIMT bank-streaming kernel.
Prints the largest difference of the denoised estimate over geometries, then times both at the cfg-1 shape (MNIST 60k, k=5)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402

dev = torch.device("cuda", 0)


def run(mod, t, x, variant):
    mod.engine(dev).ls_variant = variant
    return mod(torch.tensor([t] * x.shape[0]), x, device=dev)


worst = {"f16": 0.0, "f16x2": 0.0, "auto": 0.0}
for (H, n, ks) in () if os.environ.get("CDS_LS_DEBUG") else ((28, 300, (3, 5, 7, 9, 13, 17)), (32, 130, (3, 5, 11)), (16, 77, (3, 7, 15, 21)), (20, 64, (5, 9))):
    bank, labels = synthetic_bank(n, 1, H, nlabels=2, seed=3)
    for k in ks:
        for prec, B, t in (("f16", 1, 0.9), ("f16", 3, 0.6), ("f16x2", 2, 0.6), ("f16x2", 2, 0.15), ("auto", 2, 0.3)):
            mod = LocalScoreModule((bank, labels), kernel_size=k, batch_size=n, image_size=H,
                                   schedule=cosine_noise_schedule, precision=prec)
            beta = float(cosine_noise_schedule(torch.tensor([t]))[0])
            x = noisy_query(bank, beta, B, seed=k).to(dev)
            eng = mod.engine(dev)
            used = eng.ls_umma_supported(k, eng.passes_for(k, beta))
            s1 = run(mod, t, x, "auto")
            s0 = run(mod, t, x, "simt")
            err = float((s1 - s0).abs().max()) * beta / (1 - beta) ** 0.5
            if used:
                worst[prec] = max(worst[prec], err)
            print(f"H={H} k={k:2d} B={B} t={t} {prec}: tensor-core LS used={used} max |mu diff| = {err:.2e}", flush=True)
print(f"worst {worst}")

if os.environ.get("CDS_LS_TIME", "1") == "1":
    bank, labels = synthetic_bank(60000, 1, 28, nlabels=10, seed=0)
    mod = LocalScoreModule((bank, labels), kernel_size=5, batch_size=60000, image_size=28, schedule=cosine_noise_schedule,
                           precision="auto")
    eng = mod.engine(dev)
    sel = mod.selection(None)

    def timed(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"cfg-1 shape (MNIST 60k, k=5) {name}: {ms:.3f} ms  = {60000 * 784 * 4 / ms * 1e-6:.0f} GB/s of fp32 bank", flush=True)

    for B in ((10,) if os.environ.get("CDS_LS_DEBUG") else (1, 10)):
        for t in ((0.8,) if os.environ.get("CDS_LS_DEBUG") else (0.8, 0.2)):
            beta = cosine_noise_schedule(torch.tensor([t] * B)).to(dev, torch.float32)
            x = noisy_query(bank[:64], float(beta[0]), B, seed=1).to(dev)
            passes = eng.passes_for(5, float(beta[0]))
            timed(f"B={B} t={t} SIMT", lambda: eng.ls_partials(x, beta, 5, sel))
            timed(f"B={B} t={t} tcgen05 passes={passes}", lambda: eng.ls_partials(x, beta, 5, sel, passes=passes))
