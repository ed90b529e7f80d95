"""Debug / A-B script (GPU box): bbELS edge bands on the tensor cores vs the exact SIMT edge kernel.
Prints the largest difference of the denoised estimate over geometries, then times both on the cfg-4 shape."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalEquivBordersScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402

dev = torch.device("cuda", 0)


def run(mod, t, x, variant):
    mod.engine(dev).edge_variant = variant
    return mod(torch.tensor([t] * x.shape[0]), x, device=dev)


worst = {"f16": 0.0, "f16x2": 0.0}
for (C, H, n, ks) in ((3, 32, 70, (3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 25, 27, 31)), (1, 28, 50, (3, 7, 13, 17, 23, 27)),
                      (3, 64, 20, (3, 9, 17, 23))):
    bank, labels = synthetic_bank(n, C, H, nlabels=2, seed=3)
    for k in ks:
        for prec, B, t in (("f16", 1, 0.9), ("f16", 3, 0.75), ("f16x2", 2, 0.75), ("f16x2", 2, 0.3)):
            mod = LocalEquivBordersScoreModule((bank, labels), kernel_size=k, batch_size=16, image_size=H,
                                               schedule=cosine_noise_schedule, precision=prec)
            x = torch.randn(B, C, H, H, generator=torch.Generator().manual_seed(k)).to(dev)
            beta = float(cosine_noise_schedule(torch.tensor([t]))[0])
            eng = mod.engine(dev)
            used = eng.edge_umma_supported(k, 1 if prec == "f16" else 2)
            eng.centre_window = True
            s1 = run(mod, t, x, "auto")
            eng.centre_window = False
            s0 = run(mod, t, x, "simt")
            err = float((s1 - s0).abs().max()) * beta / (1 - beta) ** 0.5
            worst[prec] = max(worst[prec], err) if used else worst[prec]
            print(f"C={C} H={H} k={k:2d} B={B} t={t} {prec}: tensor-core edge used={used} max |mu diff| = {err:.2e}", flush=True)
print(f"worst {worst}")

if os.environ.get("CDS_EDGE_TIME", "1") == "1":
    bank, labels = synthetic_bank(50000, 3, 32, nlabels=10, seed=0)
    mod = LocalEquivBordersScoreModule((bank, labels), kernel_size=17, batch_size=64, image_size=32,
                                       schedule=cosine_noise_schedule)
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    eng = mod.engine(dev)
    sel = mod.selection(None)
    beta = cosine_noise_schedule(torch.tensor([0.9])).to(dev, torch.float32)

    def timed(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"cfg-4 shape (k=17, 50k, B=1) {name}: {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)

    timed("centre (tcgen05 ELS kernel, zero padding), all 8 query tiles", lambda: eng.umma_partials("zeros", x, beta, 17, sel, 1, a_over_beta=0.2))
    timed("centre (tcgen05 ELS kernel, zero padding), centre window = 2 query tiles",
          lambda: eng.umma_partials("zeros", x, beta, 17, sel, 1, a_over_beta=0.2, window=(8, 8, 16, 16)))
    eng.edge_variant = "simt"
    timed("edge bands, SIMT", lambda: eng.edge_partials(x, beta, 17, sel, passes=1))
    eng.edge_variant = "auto"
    timed("edge bands, tcgen05", lambda: eng.edge_partials(x, beta, 17, sel, passes=1))
    timed("corners (LS kernel)", lambda: eng.ls_partials(x, beta, 17, sel, tag="corner"))
    timed("edge bands, tcgen05, 2 passes", lambda: eng.edge_partials(x, beta, 17, sel, passes=2))
    for variant, cw in (("simt", False), ("auto", True)):
        eng.centre_window = cw
        timed(f"whole evaluation (module default precision f16x2), edge={variant}, centre window={cw}", lambda: run(mod, 0.9, x, variant))
