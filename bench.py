#!/usr/bin/env python
"""Benchmarks of the analytic score machines on B200 (one JSON line per run, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--batch B] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

Workloads = the five configurations of BASELINE.json (synthetic banks of the named shapes):

  els_cifar10_conditional (default, configs[2], the headline): ELS, 50 000 x 3x32x32 bank, class-masked, 19 evaluations with
      the kernel sizes of scales_CIFAR10_ResNet_zeros_conditional; step = B full trajectories.
  ls_mnist        (configs[0]): LS, 60 000 x 1x28x28, kernel 5, 19 evaluations, B = 10 (nsamps) trajectories per step.
  els_mnist       (configs[1]): ELS with zero padding of the query, 60 000 x 1x32x32 (the reference resizes MNIST to 32,
      src/utils/data.py:66), scales_MNIST_ResNet_zeros; step = B trajectories.
  bbels_cifar10_k17 (configs[3]): bbELS (zero padding, border-aware candidates), 50 000 x 3x32x32, kernel 17, t = 0.9;
      step = one score evaluation of B samples (the ELS/circular counterpart is timed next to it in `roofline`).
  els_celeba64_sweep (configs[4]): ELS on 50 000 x 3x64x64, step = one evaluation at each kernel size 3..17 of B samples
      per GPU (--shard replica: every rank holds the bank and its own B samples, weak scaling; --shard bank: the bank
      is split and every rank evaluates the same B samples).

N > 1 shards the bank across ranks unless --shard replica (one all-gather of partials per evaluation, fused merge).
`value` = query x train patch-pairs / s over all ranks with inputs resident in HBM; `e2e` = the same through the public
API from pinned host buffers (H2D + D2H inside the timed region).  `--impl reference` times the reference's CPU algorithm
on a bounded sample of the same workload: the real reference when $REF_DIR (or /root/reference) holds it, else
oracle/score_port.py, the port pinned against the reference-generated goldens.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "pairs/s"

WORKLOADS = {
    # name: kind, C, H, N, nlabels, conditional, scales name or fixed k, query_pad, default batch, mode, bound
    "els_cifar10_conditional": dict(kind="ELS", C=3, H=32, N=50000, nlabels=10, conditional=True,
                                    scales="CIFAR10_ResNet_zeros_conditional", pad="circular", batch=4, mode="trajectory",
                                    bound="tensor", batch_size=64,
                                    metric="ELS query x train patch-pairs/sec (CIFAR-10-shape conditional sampler)"),
    "ls_mnist": dict(kind="LS", C=1, H=28, N=60000, nlabels=10, conditional=False, k=5, pad="zeros", batch=10,
                     mode="trajectory", bound="hbm", batch_size=60000,
                     metric="LS query x train pixel-pairs/sec (MNIST-shape sampler, kernel 5)"),
    "els_mnist": dict(kind="ELS", C=1, H=32, N=60000, nlabels=10, conditional=False, scales="MNIST_ResNet_zeros",
                      pad="zeros", batch=4, mode="trajectory", bound="tensor", batch_size=64,
                      metric="ELS query x train patch-pairs/sec (MNIST-shape sampler, zero padding)"),
    "bbels_cifar10_k17": dict(kind="bbELS", C=3, H=32, N=50000, nlabels=10, conditional=False, k=17, t=0.9, pad="zeros",
                              batch=1, mode="evaluation", bound="tensor", batch_size=64,
                              metric="bbELS query x train patch-pairs/sec (CIFAR-10 shape, kernel 17)"),
    "els_celeba64_sweep": dict(kind="ELS", C=3, H=64, N=50000, nlabels=1, conditional=False, ks=[3, 5, 7, 9, 11, 13, 15, 17],
                               pad="circular", batch=32, mode="sweep", bound="tensor", batch_size=64,
                               metric="ELS query x train patch-pairs/sec (64x64x3 shape, kernel sweep 3-17)"),
}
# noise level at which a kernel size is evaluated outside a trajectory (roughly where the shipped schedules use it)
T_OF_K = {3: 0.10, 5: 0.25, 7: 0.45, 9: 0.60, 11: 0.70, 13: 0.75, 15: 0.80, 17: 0.90}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` captures of
# the same workload (a bench run cannot measure it: nothing printed under a profiler is a bench value).  Algorithmic bank bytes
# for comparison: headline 5 000 images x 48 KB of strip8 + 16 KB of norm plane = 0.32 GB per launch.
NCU_TRAFFIC = {
    "els_cifar10_conditional": {"bytes": 712.2e6, "source": "profiles/r02_els_umma_ncu_summary.md: k=17 launch of the headline "
                                "trajectory at batch 4 (705.5 MB read + 6.7 MB written; k=5: 620 MB, k=7: 592 MB); not re-measured by this run"},
    "bbels_cifar10_k17": {"bytes": 1030.4e6, "source": "profiles/r02_bbels_edge_umma_ncu_summary.md: the edge-band launch "
                          "(1 024.8 MB read + 5.7 MB written = the algorithmic 1.024 GB); not re-measured by this run"},
    "ls_mnist": {"bytes": 314.2e6, "source": "profiles/r02_ls_umma_ncu_summary.md: one B=10 evaluation (310.6 MB read + 3.5 MB written = "
                 "the unique fp16 image + fp32 norm bytes); not re-measured by this run"},
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(bf16=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def pairs_per_eval(kind, H, k, n):
    """(query pixel, candidate) pairs of one evaluation of one sample."""
    if kind == "LS" or (kind == "bbELS" and k >= H):
        return H * H * n
    if kind == "ELS":
        return H * H * n * (H - k + 1) ** 2
    d = k // 2
    i = H - 2 * d
    return n * (i ** 4 + 4 * d * i * i + 4 * d * d)


def flops_per_eval(kind, H, k, C, n):
    """Algorithmic FLOP of one evaluation of one sample: 2 per patch element and pair.  ELS pairs contract whole k*k*C
    patches; LS (and bbELS corners) a window clipped by the border; a bbELS edge pair at depth r the (r+d+1) x k rows of the
    truncated patch."""
    d = k // 2
    if kind == "ELS":
        return pairs_per_eval(kind, H, k, n) * 2 * k * k * C
    clip = [min(H - 1, y + d) - max(0, y - d) + 1 for y in range(H)]           # window extent per coordinate
    if kind == "LS" or (kind == "bbELS" and k >= H):
        return n * 2 * C * sum(clip) ** 2
    i = H - 2 * d
    edge = sum(r + d + 1 for r in range(d)) * k                                # patch elements per channel, summed over depths
    corner = sum(r + d + 1 for r in range(d)) ** 2
    return n * 2 * C * (i ** 4 * k * k + 4 * i * i * edge + 4 * corner)


def schedule_of(w):
    """[(i, k, t)] of the evaluations of one step of a workload."""
    from convolutional_diffusion_b200.scales import load_scales
    if w["mode"] == "trajectory":
        scales = load_scales(w["scales"]) if "scales" in w else [w["k"]] * 20
        n = len(scales)
        return scales, [(i, scales[i], i / n) for i in range(n - 1, 0, -1)]
    if w["mode"] == "evaluation":
        return None, [(0, w["k"], w["t"])]
    return None, [(q, k, T_OF_K[k]) for q, k in enumerate(w["ks"])]


def step_pairs_flops(w, evals, n_c):
    pairs = flops = 0
    for _, k, _ in evals:
        p = pairs_per_eval(w["kind"], w["H"], k, n_c)
        pairs += p
        flops += flops_per_eval(w["kind"], w["H"], k, w["C"], n_c)
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
def _reference_available():
    from oracle import ref_loader
    return ref_loader.available()


def cpu_sample(w, n_sub, label, seed, use_reference):
    """One step of workload `w` for ONE sample on the first n_sub bank images, on the host: the real reference when it is
    present, else the CPU port.  Returns (pairs, seconds)."""
    from oracle import score_port as sp
    from convolutional_diffusion_b200.selection import select
    from convolutional_diffusion_b200.synthetic import synthetic_bank
    bank, labels = synthetic_bank(n_sub, w["C"], w["H"], nlabels=w["nlabels"], seed=0)
    scales, evals = schedule_of(w)
    lab = label if w["conditional"] else None
    idx, logw = select(w["kind"], labels.numpy(), lab, min(w["batch_size"], n_sub), None)
    sub = bank[torch.from_numpy(idx)]
    x = torch.randn(w["C"], w["H"], w["H"], generator=torch.Generator().manual_seed(seed))
    pairs, _ = step_pairs_flops(w, evals, int(len(idx)))
    if use_reference:
        from oracle import ref_loader
        ref = ref_loader.load()
        ds = ref_loader.TensorBank(bank, labels)
        cls = {"ELS": ref.LocalEquivScoreModule, "bbELS": ref.LocalEquivBordersScoreModule, "LS": ref.LocalScoreModule}[w["kind"]]
        mod = cls(ds, kernel_size=evals[0][1], batch_size=min(w["batch_size"], n_sub), schedule=ref.cosine_noise_schedule)
        lt = None if lab is None else torch.tensor([lab])
        t0 = time.perf_counter()
        with torch.no_grad():
            if w["mode"] == "trajectory":
                ref.ScheduledScoreMachine(mod, in_channels=w["C"], imsize=w["H"], scales=scales, score_backbone=True)(
                    x[None].clone(), label=lt, device=torch.device("cpu"))
            else:
                for _, k, t in evals:
                    mod(torch.tensor([t]), x[None].clone(), label=lt, device=torch.device("cpu"), k=k)
        return pairs, time.perf_counter() - t0
    lw = torch.from_numpy(logw).float()
    t0 = time.perf_counter()
    if w["mode"] == "trajectory":
        sp.run_machine(w["kind"], x, sub, scales, lw, query_pad=w["pad"] if w["kind"] == "ELS" else None)
    else:
        for _, k, t in evals:
            beta = float(1 - math.cos(t / 1.008 * math.pi / 2) ** 2)
            sp.mu(w["kind"], x, sub, beta, k, lw, query_pad=w["pad"] if w["kind"] == "ELS" else None)
    return pairs, time.perf_counter() - t0


def cpu_sub_bank(w):
    """Bank prefix for the bounded CPU sample: about 10-30 s of work on 16 cores per step."""
    return {"els_cifar10_conditional": 2048, "ls_mnist": 60000, "els_mnist": 96, "bbels_cifar10_k17": 256,
            "els_celeba64_sweep": 8}[w["name"]]


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use_ref = _reference_available()
    n_sub = max(8, cpu_sub_bank(w) // 2)
    for i in range(args.warmup):
        cpu_sample(w, max(8, n_sub // 4), i % w["nlabels"], 1000 + i, use_ref)
    pairs = t = 0.0
    for s in range(args.steps):
        p, dt = cpu_sample(w, n_sub, s % w["nlabels"], 2000 + s, use_ref)
        pairs += p
        t += dt
    v = pairs / t
    kind = "reference" if use_ref else "port"
    sample = (f"one step of {w['name']} for b=1 on the first {n_sub} images of the same synthetic bank "
              f"({'reference modules imported from $REF_DIR' if use_ref else 'oracle/score_port.py'}, torch fp32, {cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": w["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "bank": n_sub, "batch": 1},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, w):
    import torch.distributed as dist
    import convolutional_diffusion_b200 as cd
    from convolutional_diffusion_b200.distributed import init_from_env, shutdown
    from convolutional_diffusion_b200.synthetic import synthetic_bank

    rank, world, local = init_from_env()
    dev = torch.device("cuda", local)
    B, C, H, NL = args.batch, w["C"], w["H"], w["nlabels"]
    replica = world > 1 and args.shard == "replica"
    scales, evals = schedule_of(w)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))      # torchrun pins OMP to 1 thread: bank synthesis
    bank, labels = synthetic_bank(w["N"], C, H, nlabels=NL, seed=0)
    group = dist.group.WORLD if (world > 1 and not replica) else None
    cls = {"ELS": cd.LocalEquivScoreModule, "bbELS": cd.LocalEquivBordersScoreModule, "LS": cd.LocalScoreModule}[w["kind"]]
    kw = dict(query_pad=w["pad"]) if w["kind"] == "ELS" else {}
    mod = cls((bank, labels), kernel_size=evals[0][1], batch_size=w["batch_size"], schedule=cd.cosine_noise_schedule,
              precision=args.precision, process_group=group, **kw)
    machine = cd.ScheduledScoreMachine(mod, in_channels=C, imsize=H, scales=scales, use_cuda_graph=not args.no_graph)
    eng = mod.engine(dev)
    n_per_label = [int((labels == c).sum()) for c in range(NL)] if w["conditional"] else [w["N"]] * max(NL, 1)
    del bank

    def inputs(step):
        g = torch.Generator().manual_seed(10_000 + step + (1000 * rank if replica else 0))
        return torch.randn(B, C, H, H, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    betas = {k: torch.full((B,), float(cd.cosine_noise_schedule(torch.tensor([t]))), device=dev) for _, k, t in evals}
    score_buf = torch.empty(B, C, H, H, device=dev)

    def one_step(x, lab):
        """One step on the device; returns a tensor whose read-back completes the step."""
        if w["mode"] == "trajectory":
            return machine(x, label=lab if w["conditional"] else None, device=dev)
        sel = mod.selection(lab if w["conditional"] else None)
        for _, k, t in evals:
            b = betas[k]
            eng.evaluate(w["kind"], x, b, k, sel, query_pad=mod.query_pad, mu=None, score=score_buf,
                         beta_min=float(cd.cosine_noise_schedule(torch.tensor([t]))))
        return score_buf

    nsteps_total = args.warmup + args.steps
    xs_dev = [inputs(s).to(dev) for s in range(nsteps_total)]
    xs_host = [inputs(s).pin_memory() for s in range(nsteps_total)]
    launches0 = eng.launches
    if w["mode"] == "trajectory":
        machine._forward_native(xs_dev[0], len(scales), 0 if w["conditional"] else None, dev, record=[])   # eager dry run
    elif w["mode"] == "evaluation":
        one_step(xs_dev[0], 0)
    launches_per_step = eng.launches - launches0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # L2 flush buffer (> 126 MB)

    for s in range(args.warmup):
        one_step(xs_dev[s], s % NL)
        if s == 0 and w["mode"] == "sweep":                                  # (a 64x64 sweep takes seconds: no separate dry run)
            launches_per_step = eng.launches - launches0
    if w["mode"] == "trajectory" and w["conditional"]:
        for lab in range(NL):                                                # capture every label's graph up front
            one_step(xs_dev[0], lab)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pairs = flops = 0
    nrep = world if replica else 1                                           # replicas: every rank does its own B samples
    for s in range(args.steps):
        flush.zero_()
        lab = (args.warmup + s) % NL
        ev[s][0].record()
        one_step(xs_dev[args.warmup + s], lab)
        ev[s][1].record()
        p, f = step_pairs_flops(w, evals, n_per_label[lab])
        pairs += B * p * nrep
        flops += B * f * nrep
    barrier()
    sampler.stop_flag = True
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    clk = torch.tensor([float(sampler.summary()["sm_mhz"] or 0)], dtype=torch.float64, device=dev)
    clks = [clk.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_gather(clks, clk)
    dev_ms = float(tmax.item())

    # ---- end to end through the public API: pinned host x -> device, result back to the host, every step
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        lab = (args.warmup + s) % NL
        x = xs_host[args.warmup + s].to(dev, non_blocking=True)
        one_step(x, lab).cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    # ---- dominant kernel in isolation, evaluation by evaluation
    peaks = measured_peaks()
    # (every rank runs it: with a sharded bank the recorded trajectory and the bbELS evaluation contain the all-gather)
    roof = roofline(w, args, cd, mod, machine, eng, evals, scales, xs_dev[0], dev, peaks, B)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        use_ref = _reference_available()
        n_sub = cpu_sub_bank(w)
        cpu_sample(w, max(8, n_sub // 8), 0, 1, use_ref)                       # warm the thread pool
        p, dt = cpu_sample(w, n_sub, 0, 2, use_ref)
        cpu = {"value": p / dt, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port", "seconds": dt,
               "sample": f"one step of {w['name']} for b=1 on the first {n_sub} images of the same synthetic bank "
                         f"({'reference modules from $REF_DIR' if use_ref else 'oracle/score_port.py'}, torch fp32, {cores} threads)"}

    if rank == 0:
        value = pairs / (dev_ms * 1e-3)
        par = "single" if world == 1 else (f"replicas x{world} (bank resident on every rank, {B} samples each)" if replica
                                           else f"bank-shard x{world}")
        out = {
            "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak" if replica else "strong",
            "vs_baseline": None,
            "dtype": ("f32" if w["kind"] == "LS" else
                      "f16 operands (hi+lo query split where (a/beta)*k*sqrt(C) > 20), f32 accumulate/softmax"
                      if args.precision != "f16" else "f16"),
            "data": "synthetic",
            "config": {"workload": w["name"], "bank": w["N"], "image": [C, H, H], "kind": w["kind"],
                       "scales": w.get("scales", w.get("ks", w.get("k"))), "evals_per_step": len(evals), "batch": B,
                       "precision": args.precision, "parallelism": par,
                       "l2": "256 MB buffer written between timed steps; the bank streamed per evaluation is larger than L2"},
            "samples_per_s": B * nrep * args.steps / (dev_ms * 1e-3),
            "tflops_algorithmic": flops / (dev_ms * 1e-3) * 1e-12,
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "samples_per_s": B * nrep * args.steps / e2e_s,
                    "h2d_bytes_per_step": B * C * H * H * 4, "d2h_bytes_per_step": B * C * H * H * 4},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": dict(sampler.summary(), per_rank_sm_mhz=[float(c.item()) for c in clks]),
            "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
        sys.stdout.flush()
    machine.release_graphs()
    shutdown()


def roofline(w, args, cd, mod, machine, eng, evals, scales, x0, dev, peaks, B):
    """Per-evaluation CUDA-event timing of the dominant kernel of the workload on the x of its own step."""
    lab = 0 if w["conditional"] else None
    sel = mod.selection(lab)
    n_c = sel[2]
    if w["mode"] == "trajectory":
        _, rec = machine.trajectory(x0, label=None if lab is None else torch.tensor([lab]), device=dev)
        x_at = {r["i"]: r["x"].contiguous() for r in rec}
    else:
        x_at = {i: x0 for i, _, _ in evals}
    per_eval = {}
    for i, k, t in evals:
        beta_val = float(cd.cosine_noise_schedule(torch.tensor([t])))
        beta = torch.full((B,), beta_val, device=dev)
        x = x_at[i]
        if w["kind"] == "LS":
            passes = eng.passes_for(k, beta_val)
            fn = lambda: eng.ls_partials(x, beta, k, sel, passes=passes)
        elif w["kind"] == "ELS":
            passes = eng.passes_for(k, beta_val)
            aob = eng._a_over_beta(beta_val)
            fn = lambda: eng.umma_partials(w["pad"], x, beta, k, sel, passes, a_over_beta=aob)
        else:
            mu = torch.empty_like(x)
            fn = lambda: eng.evaluate("bbELS", x, beta, k, sel, mu=mu, beta_min=beta_val)
        heavy = w["mode"] == "sweep"                          # a 64x64 evaluation takes seconds and the sweeps before this
        for _ in range(0 if heavy else 2):                    # point have warmed every size: one timed launch
            fn()
        reps = 1 if heavy else 3
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize()
        per_eval.setdefault(k, []).append(a.elapsed_time(b_) / reps)
    per_k, tot_ms, tot_fl, tot_bytes = {}, 0.0, 0.0, 0.0
    for k in sorted(per_eval):
        count = len(per_eval[k])
        ms = sum(per_eval[k]) / count
        p = B * pairs_per_eval(w["kind"], w["H"], k, n_c)
        fl = B * flops_per_eval(w["kind"], w["H"], k, w["C"], n_c)
        per_k[str(k)] = {"ms": round(ms, 4), "pairs_per_s": p / ms * 1e3, "tflops": fl / ms * 1e-9, "evals": count}
        tot_ms += ms * count
        tot_fl += fl * count
        tot_bytes += n_c * w["C"] * w["H"] * w["H"] * eng.bank.ls_bytes_per_pixel() * count
    if w["bound"] == "hbm":
        achieved = tot_bytes / tot_ms * 1e-6                 # GB/s
        return {"bound": "hbm", "kernel": "ls_umma_kernel" if eng.ls_umma_supported(w["k"], 1) else "ls_rows_kernel", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": achieved / peaks["hbm"], "peak_source": peaks["source"],
                "traffic": NCU_TRAFFIC.get(w["name"], {}).get("bytes"), "traffic_source": NCU_TRAFFIC.get(w["name"], {}).get("source"),
                "note": f"algorithmic bytes = selected images x C*H*W x {eng.bank.ls_bytes_per_pixel()} B (the bank streamed once per "
                        f"evaluation for all {B} samples) / CUDA-event kernel time", "per_k": per_k}
    achieved = tot_fl / tot_ms * 1e-9
    out = {"bound": "tensor", "kernel": "els_umma_kernel" if w["kind"] == "ELS" else "bbELS: els_umma_kernel (centre window) + bbels_edge_umma_kernel (edge bands) + ls_rows_kernel (corners)",
           "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
           "frac_sustained": achieved / peaks["bf16_sustained"], "peak_source": peaks["source"],
           "traffic": NCU_TRAFFIC.get(w["name"], {}).get("bytes"),
           "traffic_source": NCU_TRAFFIC.get(w["name"], {}).get("source"),
           "note": "algorithmic 2 FLOP per patch element and (query, patch) pair (2*k*k*C for ELS; truncated patches counted as such for bbELS edges/corners), FLOP-weighted over the evaluations of one step, each on the "
                   "x of its own step; CUDA events around the launches on the launching stream", "per_k": per_k}
    if w["name"] == "bbels_cifar10_k17":                     # the ELS / circular counterpart of configs[3]
        k, t = w["k"], w["t"]
        beta_val = float(cd.cosine_noise_schedule(torch.tensor([t])))
        beta = torch.full((B,), beta_val, device=dev)
        mu = torch.empty_like(x0)
        for _ in range(2):
            eng.evaluate("ELS", x0, beta, k, sel, query_pad="circular", mu=mu, beta_min=beta_val)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            eng.evaluate("ELS", x0, beta, k, sel, query_pad="circular", mu=mu, beta_min=beta_val)
        b_.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / 3
        p = B * pairs_per_eval("ELS", w["H"], k, n_c)
        out["els_circular_counterpart"] = {"ms": ms, "pairs_per_s": p / ms * 1e3, "tflops": p * 2 * k * k * w["C"] / ms * 1e-9}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="els_cifar10_conditional", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--shard", default="bank", choices=["bank", "replica"])
    ap.add_argument("--bank", type=int, default=None, help="bank size override (smoke runs; the named size is the default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "f16", "f16x2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (for ncu launch lists)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload], name=args.workload)
    if args.batch is None:
        args.batch = w["batch"]
    if args.bank is not None:
        w["N"] = args.bank
    if args.impl == "ours" and w["mode"] != "sweep":
        args.warmup = max(args.warmup, 3)                     # (the 64x64 sweep takes seconds per step: --warmup is honoured)
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
