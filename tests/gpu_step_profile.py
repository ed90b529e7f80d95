"""Per-evaluation timing of the tensor-core ELS kernel ALONG A REAL TRAJECTORY of the headline workload (50k CIFAR-shape
bank, class-masked, batch 4): x at step i is what the sampler actually feeds the kernel, so the skip rates are the real
ones (bench.py's per-k roofline numbers use x = randn at every noise level).  Not a pytest file.

    python tests/gpu_step_profile.py                       # timings only
    CDS_LIB_PATH=.../libcdscore_prof.so python tests/gpu_step_profile.py   # + chunk skip counters (profile build)
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutional_diffusion_b200 as cd  # noqa: E402
from convolutional_diffusion_b200 import _lib  # noqa: E402
from convolutional_diffusion_b200.scales import load_scales  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402


def main():
    B = int(os.environ.get("CDS_BATCH", "4"))
    label = 0
    scales = load_scales("CIFAR10_ResNet_zeros_conditional")
    bank, labels = synthetic_bank(50000, 3, 32, seed=0)
    mod = cd.LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cd.cosine_noise_schedule,
                                   precision="auto")
    machine = cd.ScheduledScoreMachine(mod, in_channels=3, imsize=32, scales=scales)
    eng = mod.engine("cuda")
    sel = mod.selection(label)
    lib = _lib.load()
    have_counters = hasattr(lib, "cds_debug_els_counters")
    x0 = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(10_000)).cuda()
    cache = os.environ.get("CDS_TRAJ_CACHE")        # profiling switches corrupt the trajectory: take x from a clean run
    if cache and os.path.isfile(cache):
        rec = torch.load(cache)
        for r in rec:
            r["x"] = r["x"].cuda()
    else:
        _, rec = machine.trajectory(x0, label=torch.tensor([label]), device="cuda")
        if cache:
            torch.save([dict(i=r["i"], k=r["k"], beta=r["beta"], x=r["x"].cpu()) for r in rec], cache)
    only = os.environ.get("CDS_STEPS")
    if only:
        keep = {int(v) for v in only.split(",")}
        rec = [r for r in rec if r["i"] in keep]
    print(f"# n_sel={sel[2]} B={B} counters={'yes' if have_counters else 'no'}")
    print("# i k beta a/beta passes ms pairs/s TFLOP/s chunk_skip% warp_tile_skip% newmax%")
    tot = 0.0
    for r in rec:
        k, beta_val, x = r["k"], r["beta"], r["x"].contiguous()
        beta = torch.full((B,), beta_val, device="cuda")
        passes = eng.passes_for(k, beta_val)
        for _ in range(2):
            eng.umma_partials("circular", x, beta, k, sel, passes, a_over_beta=eng._a_over_beta(beta_val))
        torch.cuda.synchronize()
        cnt = (ctypes.c_ulonglong * 16)()
        if have_counters:
            lib.cds_debug_els_counters(cnt)       # clear
        reps = 3
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.umma_partials("circular", x, beta, k, sel, passes, a_over_beta=eng._a_over_beta(beta_val))
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tot += ms
        pairs = B * 1024 * sel[2] * (33 - k) ** 2
        extra = ""
        if have_counters:
            lib.cds_debug_els_counters(cnt)
            c = [int(v) for v in cnt]
            extra = (f" {100.0 * c[1] / max(c[0], 1):.1f} {100.0 * c[4] / max(c[3], 1):.1f} {100.0 * c[2] / max(c[0], 1):.2f}"
                     f" exact%={100.0 * c[5] / max(c[0], 1):.3f} drains={c[6]}")
            if c[10] and c[13]:
                extra += (f" | clk/tile MMA wait={c[8] / c[10]:.0f} issue={c[9] / c[10]:.0f}  EPI wait={c[11] / c[13]:.0f} work={c[12] / c[13]:.0f}"
                          f" [ld={c[14] / c[13]:.0f} max+classify={c[15] / c[13]:.0f} weights={c[2] / c[13]:.0f} st={c[3] / c[13]:.0f} tail={c[4] / c[13]:.0f}]")
        print(f"{r['i']:2d} {k:2d} {beta_val:.5f} {((1 - beta_val) ** 0.5) / beta_val:7.2f} {passes} {ms:7.3f} "
              f"{pairs / ms * 1e3:.3e} {pairs * 2 * k * k * 3 / ms * 1e-9:7.1f}{extra}")
    print(f"# sum of the 19 evaluations: {tot:.2f} ms")


if __name__ == "__main__":
    main()
