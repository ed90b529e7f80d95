"""GPU diagnostic (not a pytest file): raw tcgen05 dot products vs numpy, per geometry.
Run on the B200 box:  python tests/gpu_debug_umma.py > gpurun_out/umma_debug.log"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import PatchBank, ScoreEngine  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402
from numpy.lib.stride_tricks import sliding_window_view  # noqa: E402


def dots_ref(x, img, k, pad):
    c, h, w = x.shape
    d = k // 2
    xp = np.pad(x, ((0, 0), (d, d), (d, d)), mode="wrap" if pad == "circular" else "constant")
    q = sliding_window_view(xp, (k, k), axis=(1, 2)).transpose(1, 2, 0, 3, 4).reshape(h * w, -1)
    p = sliding_window_view(img, (k, k), axis=(1, 2)).transpose(1, 2, 0, 3, 4).reshape(-1, q.shape[1])
    return q @ p.T


def main():
    torch.manual_seed(0)
    cases = [(3, 32, 3, 2, "circular"), (3, 32, 17, 2, "circular"), (1, 32, 5, 2, "zeros"), (3, 32, 9, 1, "circular"),
             (3, 16, 5, 2, "circular"), (1, 28, 7, 2, "circular"), (3, 32, 11, 2, "zeros"), (3, 32, 15, 2, "circular")]
    for (C, H, k, passes, pad) in cases:
        imgs, labels = synthetic_bank(6, C, H, seed=1)
        bank = PatchBank(imgs, labels, device="cuda")
        eng = ScoreEngine(bank, precision="f16x2" if passes == 2 else "f16")
        eng.els_variant = os.environ.get("CDS_VARIANT", "pv")
        ok = eng.umma_supported(k, passes)
        print(f"C={C} H={H} k={k} passes={passes} pad={pad} supported={ok}", flush=True)
        if not ok:
            continue
        x = torch.randn(2, C, H, H, device="cuda")
        beta = torch.tensor([0.5, 0.2], device="cuda")
        sel = bank.selection("ELS", None, 64, None)
        P = (H - k + 1) ** 2
        dbg = torch.full((2, H * H, P), float("nan"), device="cuda")
        t0 = time.time()
        eng.umma_partials(pad, x, beta, k, sel, passes, dbg=dbg)
        torch.cuda.synchronize()
        print(f"   kernel returned in {time.time() - t0:.3f}s", flush=True)
        for b in range(2):
            ref = dots_ref(x[b].cpu().double().numpy(), imgs[0].double().numpy(), k, pad)
            # the accumulator also carries the norm term: scale*(q.p) - (a*scale/2)|p|^2  (reported / scale)
            pn = bank.patch_norms(k).view(-1, P)[0].cpu().double().numpy()
            ref = ref - 0.5 * np.sqrt(1 - float(beta[b])) * pn[None, :]
            got = dbg[b].cpu().double().numpy()
            err = np.abs(got - ref)
            print(f"   b={b} dots: nan={np.isnan(got).sum()} max|err|={np.nanmax(err):.3e} ref_rms={ref.std():.3f}",
                  flush=True)
            if np.nanmax(err) > 0.05:
                bad = np.argwhere(err > 0.05)
                print("   first bad (query, cand):", bad[:8].tolist(), "n_bad", len(bad), "of", err.size)
    print("done")


if __name__ == "__main__":
    main()
