#!/usr/bin/env python
"""Headline benchmark: the ELS CIFAR-10-shape conditional sampler (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

One *step* = one pass of the hot path over one batch: B full trajectories (19 score evaluations each, kernel
sizes from scales_CIFAR10_ResNet_zeros_conditional, class-masked 50 000-image synthetic bank, one class label per
step as in scripts/els_script.py:194) through ScheduledScoreMachine.  N > 1 shards the bank across ranks
(strong scaling: same total work, every rank reduces its slice, one all-gather of partials per evaluation).

Prints ONE JSON line (rank 0).  `value` = query x train patch-pairs / s over all ranks with inputs resident in
HBM; `e2e` = the same through the public API from pinned host buffers (H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm (oracle/score_port.py, the port pinned against the
reference-generated goldens; the Python reference itself cannot travel to the GPU box) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ELS query x train patch-pairs/sec (CIFAR-10-shape conditional sampler)"
UNIT = "pairs/s"
SCALES_NAME = "CIFAR10_ResNet_zeros_conditional"
H = W = 32
C = 3
N_BANK = 50000
NLABELS = 10


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(bf16=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def pairs_per_eval(k, n_c):
    return H * W * n_c * (H - k + 1) * (W - k + 1)


def trajectory_pairs_flops(scales, n_c):
    pairs = flops = 0
    for i in range(len(scales) - 1, 0, -1):
        k = scales[i]
        p = pairs_per_eval(k, n_c)
        pairs += p
        flops += p * 2 * k * k * C
    return pairs, flops


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_sample(n_sub, label, scales, seed):
    """One 19-evaluation conditional trajectory of the CPU port on the first n_sub bank images."""
    from oracle import score_port as sp
    from convolutional_diffusion_b200.synthetic import synthetic_bank
    bank, labels = synthetic_bank(n_sub, C, H, nlabels=NLABELS, seed=0)
    sel = (labels == label).nonzero()[:, 0]
    sub = bank[sel]
    # ELS per-batch mean quirk (idealscore.py:470): batch_size 64 -> weight 1/n_b per image of a batch
    from convolutional_diffusion_b200.selection import select
    _, logw = select("ELS", labels.numpy(), label, 64, None)
    x = torch.randn(C, H, W, generator=torch.Generator().manual_seed(seed))
    t0 = time.perf_counter()
    sp.run_machine("ELS", x, sub, scales, torch.from_numpy(logw).float())
    dt = time.perf_counter() - t0
    pairs, _ = trajectory_pairs_flops(scales, int(sel.numel()))
    return pairs, dt


def run_reference(args, scales):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_sub = 1024
    for w in range(args.warmup):
        cpu_sample(n_sub, w % NLABELS, scales, 1000 + w)
    pairs = 0
    t = 0.0
    for s in range(args.steps):
        p, dt = cpu_sample(n_sub, s % NLABELS, scales, 2000 + s)
        pairs += p
        t += dt
    v = pairs / t
    sample = f"19-eval ELS conditional trajectory, b=1, first {n_sub} images of the same synthetic bank (~{n_sub // NLABELS} in class)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "els_cifar10_conditional", "scales": SCALES_NAME, "bank": n_sub, "batch": 1},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, scales):
    import torch.distributed as dist
    from convolutional_diffusion_b200 import LocalEquivScoreModule, ScheduledScoreMachine, cosine_noise_schedule
    from convolutional_diffusion_b200.distributed import init_from_env
    from convolutional_diffusion_b200.synthetic import synthetic_bank

    rank, world, local = init_from_env()
    dev = torch.device("cuda", local)
    B = args.batch
    bank, labels = synthetic_bank(N_BANK, C, H, nlabels=NLABELS, seed=0)
    group = dist.group.WORLD if world > 1 else None
    mod = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule,
                                precision=args.precision, process_group=group)
    machine = ScheduledScoreMachine(mod, in_channels=C, imsize=H, scales=scales, use_cuda_graph=not args.no_graph)
    eng = mod.engine(dev)
    n_per_label = [int((labels == c).sum()) for c in range(NLABELS)]

    def inputs(step):
        g = torch.Generator().manual_seed(10_000 + step)
        return torch.randn(B, C, H, W, generator=g), step % NLABELS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # resident inputs for the device-timed loop, pinned host inputs for the end-to-end loop
    nsteps_total = args.warmup + args.steps
    xs_dev = [inputs(s)[0].to(dev) for s in range(nsteps_total)]
    xs_host = [inputs(s)[0].pin_memory() for s in range(nsteps_total)]

    launches0 = eng.launches
    machine._forward_native(xs_dev[0], len(scales), 0, dev, record=[])      # eager dry run: counts the launches
    launches_per_traj = eng.launches - launches0
    # L2 flush buffer (> 126 MB) written between timed steps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    for s in range(args.warmup):
        machine(xs_dev[s], label=s % NLABELS, device=dev)
    for lab in range(NLABELS):                                               # capture every label's graph up front
        machine(xs_dev[0], label=lab, device=dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pairs = flops = 0
    for s in range(args.steps):
        flush.zero_()
        lab = (args.warmup + s) % NLABELS
        ev[s][0].record()
        machine(xs_dev[args.warmup + s], label=lab, device=dev)
        ev[s][1].record()
        p, f = trajectory_pairs_flops(scales, n_per_label[lab])
        pairs += B * p
        flops += B * f
    barrier()
    sampler.stop_flag = True
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())

    # ---- end to end through the public API: pinned host x -> device, result back to the host, every step
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        lab = (args.warmup + s) % NLABELS
        x = xs_host[args.warmup + s].to(dev, non_blocking=True)
        out = machine(x, label=torch.tensor([lab]), device=dev)
        out.cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    # ---- dominant kernel in isolation: the tcgen05 partials kernel for each of the 19 evaluations
    peaks = measured_peaks()
    roof = None
    per_k = {}
    if rank == 0:
        lab = 0
        sel = mod.selection(lab)
        n_c_local = sel[2]
        # every evaluation at its own noise level AND on the x the sampler actually feeds it at that step (an eager
        # trajectory records them): the pass count depends on the noise level, the fraction of chunks that carry no weight
        # on how close x is to the bank -- x = randn at a low noise level would be far cheaper than the real step
        _, rec = machine.trajectory(xs_dev[0], label=torch.tensor([lab]), device=dev)
        x_at = {r["i"]: r["x"].contiguous() for r in rec}
        tot_ms = tot_fl = 0.0
        per_eval = {}
        for i in range(1, len(scales)):
            k = scales[i]
            x = x_at[i]
            beta_val = float(cosine_noise_schedule(torch.tensor([i / len(scales)])))
            beta = torch.full((B,), beta_val, device=dev)
            passes = eng.passes_for(k, beta_val)
            aob = eng._a_over_beta(beta_val)
            for _ in range(2):
                eng.umma_partials("circular", x, beta, k, sel, passes, a_over_beta=aob)
            reps = 3
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                eng.umma_partials("circular", x, beta, k, sel, passes, a_over_beta=aob)
            b_.record()
            torch.cuda.synchronize()
            per_eval.setdefault(k, []).append((a.elapsed_time(b_) / reps, passes))
        for k in sorted(per_eval):
            count = len(per_eval[k])
            ms = sum(m for m, _ in per_eval[k]) / count
            p = B * pairs_per_eval(k, n_c_local)
            fl = p * 2 * k * k * C
            per_k[str(k)] = {"ms": round(ms, 4), "pairs_per_s": p / ms * 1e3, "tflops": fl / ms * 1e-9,
                             "passes": sorted(set(q for _, q in per_eval[k])), "evals": count}
            tot_ms += ms * count
            tot_fl += fl * count
        achieved = tot_fl / tot_ms * 1e-9
        roof = {"bound": "tensor", "kernel": "els_umma_kernel", "achieved": achieved, "peak": peaks["bf16"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16"], "frac_sustained": achieved / peaks["bf16_sustained"],
                "peak_source": peaks["source"],
                # DRAM bytes per launch of this kernel (ncu --set full, profiles/r01g_els_umma_ncu_summary.md: dram read+write
                # at batch 4, class 0, k=17: 712 MB; k=5: 621 MB, k=11: 917 MB); algorithmic = class sub-bank strip8 + norm
                # plane + the rows8 rows of the mixed K layout once = 446 MB at k=17 (324 MB at k <= 7)
                "traffic": 712.0e6, "traffic_algorithmic": 446.0e6,
                "traffic_source": "ncu --set full capture of round 1 at batch 4, class 0, k=17 (profiles/r01g_els_umma_ncu_summary.md); "
                                  "not re-measured by this run",
                "note": "algorithmic 2*k*k*C FLOP per (query, patch) pair, FLOP-weighted over the 19 evaluations of one trajectory, each "
                        "on the x of its own step; "
                        "trajectory; CUDA events around the kernel launches on the launching stream", "per_k": per_k}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_sub = 2048
        cpu_sample(256, 0, scales, 1)                                          # warm the thread pool
        p, dt = cpu_sample(n_sub, 0, scales, 2)
        cpu = {"value": p / dt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
               "sample": f"one 19-eval ELS conditional trajectory, b=1, first {n_sub} images of the same synthetic bank "
                         f"(oracle/score_port.py, torch fp32, {cores} threads)"}

    if rank == 0:
        value = pairs / (dev_ms * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands (hi+lo query split where (a/beta)*k*sqrt(C) > 20), f32 accumulate/softmax" if args.precision != "f16" else "f16",
            "data": "synthetic",
            "config": {"workload": "els_cifar10_conditional", "bank": N_BANK, "image": [C, H, W], "scales": SCALES_NAME,
                       "evals_per_trajectory": len(scales) - 1, "batch": B, "precision": args.precision,
                       "parallelism": f"bank-shard x{world}" if world > 1 else "single",
                       "l2": "256 MB buffer written between timed steps; class sub-bank streamed per evaluation is 235 MB > L2"},
            "samples_per_s": B * args.steps / (dev_ms * 1e-3),
            "tflops_algorithmic": flops / (dev_ms * 1e-3) * 1e-12,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "samples_per_s": B * args.steps / e2e_s,
                    "h2d_bytes_per_step": B * C * H * W * 4, "d2h_bytes_per_step": B * C * H * W * 4},
            "gpu_launches": launches_per_traj * args.steps,
            "clocks": sampler.summary(),
            "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        # no collective after the timed region: tearing down a NCCL communicator that is referenced by captured CUDA
        # graphs can block, so every rank just drains its own stream and leaves
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "f16x2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from convolutional_diffusion_b200.scales import load_scales
    scales = load_scales(SCALES_NAME)
    if args.impl == "reference":
        run_reference(args, scales)
    else:
        run_ours(args, scales)


if __name__ == "__main__":
    main()
