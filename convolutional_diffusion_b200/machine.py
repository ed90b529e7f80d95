"""ScheduledScoreMachine: the deterministic DDIM reverse loop that drives the score modules.

Mirrors `/root/reference/src/utils/idealscore.py:55-124` (constructor, forward, sample):
    for i = nsteps-1 ... 1:  t = i/nsteps, k = scales[i], s = backbone(t, x, label, device, k),
                             eps = -sqrt(beta_t) s,  x <- sqrt(a'/a) x + (sqrt(b') - sqrt(a'/a) sqrt(b)) eps
With one of this package's score modules as backbone the loop runs entirely on the device: each step is
[partials kernel(s) -> merge -> epilogue -> fused DDIM update in terms of the denoised estimate
x <- sqrt(b'/b) x + (sqrt(a') - sqrt(b'/b) sqrt(a)) mu], and the whole trajectory is captured once per
(scales, batch, label) in a CUDA graph and replayed.  Any other callable backbone takes the generic loop.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .modules import _ScoreModuleBase, _as_label, _label_groups, cosine_noise_schedule


def ddim_coefficients(nsteps, schedule=cosine_noise_schedule):
    """Per step i = nsteps-1 ... 1: (i, beta_t, c_x, c_mu) of  x <- c_x x + c_mu mu  (idealscore.py:88-116)."""
    out = []
    for i in range(nsteps - 1, 0, -1):
        t = torch.tensor([i / nsteps], dtype=torch.float32)
        bt = float(schedule(t))
        bp = max(float(schedule(t - 1 / nsteps)), 0.0)
        r = math.sqrt(bp / bt)
        out.append((i, bt, r, math.sqrt(1.0 - bp) - r * math.sqrt(1.0 - bt)))
    return out


def ddpm_coefficients(nsteps, schedule=cosine_noise_schedule, first_step=None):
    """The stochastic (DDPM) branch of the reference sampler, `DDIM.sample(ddpm=True)` (src/models.py:48-64):
        sigma = sqrt(b'/b) sqrt(1 - a/a'),   x <- sqrt(a') (x - sqrt(b) eps)/sqrt(a) + sqrt(1 - a' - sigma^2) eps + sigma z
    written in terms of the denoised estimate mu = (x - sqrt(b) eps)/sqrt(a):
        x <- c_x x + c_mu mu + sigma z,   c_x = sqrt(b' - sigma^2)/sqrt(b),   c_mu = sqrt(a') - c_x sqrt(a).
    Per step i = first_step ... 1: (i, beta_t, c_x, c_mu, sigma).  first_step defaults to nsteps-1 (the loop of
    ScheduledScoreMachine.forward, idealscore.py:88); DDIM.sample itself starts at i = nsteps."""
    out = []
    for i in range(nsteps - 1 if first_step is None else int(first_step), 0, -1):
        t = torch.tensor([i / nsteps], dtype=torch.float32)
        bt = float(schedule(t))
        bp = max(float(schedule(t - 1 / nsteps)), 0.0)
        at, ap = 1.0 - bt, 1.0 - bp
        sigma = math.sqrt(bp / bt) * math.sqrt(max(1.0 - at / ap, 0.0))
        ce = math.sqrt(max(bp - sigma * sigma, 0.0))
        cx = ce / math.sqrt(bt)
        out.append((i, bt, cx, math.sqrt(ap) - cx * math.sqrt(at), sigma))
    return out


class ScheduledScoreMachine(nn.Module):
    def __init__(self, backbone, in_channels=3, imsize=32, default_time_steps=20,
                 noise_schedule=cosine_noise_schedule, score_backbone=True, scales=None, use_cuda_graph=True,
                 **kwargs):
        super().__init__()
        self.backbone = backbone
        self.default_time_steps = default_time_steps
        self.noise_schedule = noise_schedule
        self.in_channels = in_channels
        self.imsize = imsize
        self.score_backbone = score_backbone
        # plain ints: calibrate()['median'] and torch.load results arrive as tensors, whose 0-d elements hash by identity
        # (a graph-cache miss on every call) -- normalise once
        self.scales = None if scales is None else [int(s) for s in (scales.tolist() if torch.is_tensor(scales) else scales)]
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self.max_cached_graphs = 32          # (nsteps, B, label, scales) keys; oldest evicted beyond this

    # ------------------------------------------------------------------------------------------
    def forward(self, x, nsteps=None, label=None, device=None, visualize=False, ddpm=False, noise=None, seed=None,
                first_step=None):
        """Deterministic DDIM trajectory as in the reference; ddpm=True takes the stochastic branch of the reference's
        sampler (`DDIM.sample(ddpm=True)`, src/models.py:48-64) instead.  Its Gaussian noise is drawn on the device by a
        Philox4x32-10 stream keyed by `seed` (default: one draw from torch's global RNG), inside the captured CUDA graph;
        `noise` (a sequence of [B,C,H,W] tensors, one per step) injects given noise instead -- that is how the parity tests
        replay a reference run.  first_step (DDPM only) starts the loop at another step, e.g. nsteps as DDIM.sample does."""
        if device is None:
            device = torch.device("cuda")
        if nsteps is None:
            nsteps = self.default_time_steps if self.scales is None else len(self.scales)
        native = isinstance(self.backbone, _ScoreModuleBase) and self.score_backbone \
            and self.backbone.schedule is self.noise_schedule
        if ddpm and not native:
            raise NotImplementedError("the stochastic sampler runs on this package's score modules only")
        sto = None
        if ddpm:
            sto = dict(noise=noise, seed=int(torch.empty((), dtype=torch.int64).random_().item()) if seed is None else int(seed),
                       first_step=first_step)
        groups = _label_groups(label, x.shape[0]) if native else None
        if groups is not None:                   # per-sample labels: one trajectory per distinct label
            out = None
            for lab_g, rows in groups.items():
                sub = None if sto is None else dict(sto, noise=None if noise is None else [z[rows] for z in noise])
                res = self._forward_native(x[rows], nsteps, lab_g, torch.device(device), sto=sub)
                if out is None:
                    out = torch.empty((x.shape[0],) + tuple(res.shape[1:]), dtype=res.dtype, device=res.device)
                out[torch.as_tensor(rows, device=res.device)] = res
            return out
        if native:
            return self._forward_native(x, nsteps, _as_label(label), torch.device(device), sto=sto)
        return self._forward_generic(x, nsteps, label, device)

    def sample(self, nsteps=None, label=None, device=None):
        if device is None:
            device = torch.device("cuda")
        x = torch.randn(1, self.in_channels, self.imsize, self.imsize, device=device)
        return self(x.clone(), nsteps=nsteps, label=label, device=device)

    # ---- generic loop: any callable backbone, same arithmetic as the reference --------------------
    def _forward_generic(self, x, nsteps, label, device):
        x = x.clone()
        for i in range(nsteps - 1, 0, -1):
            bsz = x.shape[0]
            t = i * torch.ones(bsz) / nsteps
            beta_t = self.noise_schedule(t).to(device)
            k = None if self.scales is None else self.scales[i]
            if label is not None:
                eps = self.backbone(t, x, label=label, device=device, k=k)
            else:
                eps = self.backbone(t, x, device=device, k=k)
            if self.score_backbone:
                eps = eps * (-beta_t ** 0.5)[:, None, None, None]
            alpha_t = 1 - beta_t
            beta_prev = self.noise_schedule(t - 1 / nsteps).to(device)
            alpha_prev = 1 - beta_prev
            ratio = ((alpha_prev / alpha_t) ** 0.5)[:, None, None, None]
            x = x * ratio + (beta_prev[:, None, None, None] ** 0.5 - ratio * beta_t[:, None, None, None] ** 0.5) * eps
        return x

    # ---- native loop: device resident, graph captured ----------------------------------------------
    def _plan(self, nsteps, B, device, sto=None):
        mod = self.backbone
        steps = []
        if sto is None:
            coeffs = [(i, bt, cx, cmu, None) for i, bt, cx, cmu in ddim_coefficients(nsteps, self.noise_schedule)]
        else:
            coeffs = ddpm_coefficients(nsteps, self.noise_schedule, sto.get("first_step"))
        for i, bt, cx, cmu, sigma in coeffs:
            k = mod.kernel_size if self.scales is None else int(self.scales[min(i, len(self.scales) - 1)])
            st = dict(i=i, k=k, beta=bt,
                      beta_dev=torch.full((B,), bt, dtype=torch.float32, device=device),
                      cx=torch.full((B,), cx, dtype=torch.float32, device=device),
                      cmu=torch.full((B,), cmu, dtype=torch.float32, device=device), index=i)
            if sigma is not None:
                st["sigma"] = torch.full((B,), sigma, dtype=torch.float32, device=device)
            steps.append(st)
        return steps

    def _run_steps(self, eng, steps, x, mu, sel, sel_ls, record=None, label=None, reselect=False):
        """reselect: the visiting order is shuffled (LS, shuffle=True, or the bbELS -> LS delegation with
        batch_size < N).  The reference opens a fresh DataLoader iterator per score evaluation (idealscore.py:184,430,521),
        i.e. a new permutation -- and two global RNG draws -- at every step, so the selection is redrawn per step
        (eager path only; a captured graph would freeze one permutation)."""
        mod = self.backbone
        for st in steps:
            if reselect:
                if mod.kind == "bbELS":
                    if st["k"] >= eng.bank.H:
                        sel_ls = mod.selection(label, kind="LS")
                    elif mod.shuffle:
                        sel = mod.selection(label)
                else:
                    sel = mod.selection(label)
            x_before = x.clone() if record is not None else None
            # one fused tail per evaluation: merge of the partials, mu, and the sampler update of x in place
            eng.evaluate(mod.kind, x, st["beta_dev"], st["k"], sel, query_pad=mod.query_pad, mu=mu, score=None,
                         beta_min=st["beta"], sel_ls=sel_ls, step=st)
            if record is not None:
                record.append(dict(i=st["i"], k=st["k"], beta=st["beta"], x=x_before, mu=mu.clone()))

    def _forward_native(self, x, nsteps, label, device, record=None, sto=None):
        mod = self.backbone
        eng = mod.engine(device)
        B = x.shape[0]
        needs_ls = mod.kind == "bbELS" and any(
            (mod.kernel_size if self.scales is None else int(s)) >= eng.bank.H
            for s in (self.scales[1:nsteps] if self.scales is not None else [mod.kernel_size]))
        shuffled = mod.shuffle or (mod.kind == "LS" and mod._ls_shuffles()) or needs_ls and mod._ls_shuffles()
        if shuffled:
            # drawn per step inside _run_steps (one permutation = two global RNG draws per evaluation, as in the
            # reference); only the deterministic part is resolved here
            sel = mod.selection(label) if (mod.kind == "bbELS" and not mod.shuffle) else None
            sel_ls = None
        else:
            sel = mod.selection(label)
            sel_ls = mod.selection(label, kind="LS") if needs_ls else None
        injected = sto is not None and sto.get("noise") is not None
        key = (nsteps, B, label, tuple(self.scales) if self.scales is not None else None,
               None if sto is None else ("ddpm", sto.get("first_step")))
        with torch.cuda.device(eng.device):
            if record is not None or not self.use_cuda_graph or shuffled or injected:
                xw = x.to(eng.device, torch.float32).clone().contiguous()
                mu = torch.empty_like(xw)
                steps = self._plan(nsteps, B, eng.device, sto)
                if sto is not None:
                    so = torch.tensor([sto["seed"], 0], dtype=torch.int64, device=eng.device)
                    for q, st in enumerate(steps):
                        st["seed_offset"] = so
                        if injected:
                            st["noise"] = sto["noise"][q].to(eng.device, torch.float32).contiguous()
                self._run_steps(eng, steps, xw, mu, sel, sel_ls, record, label=label, reselect=bool(shuffled))
                return xw
            if key not in self._graphs:
                steps = self._plan(nsteps, B, eng.device, sto)
                so = None
                if sto is not None:          # the seed lives in a device buffer: rewritten before every replay
                    so = torch.zeros(2, dtype=torch.int64, device=eng.device)
                    for st in steps:
                        st["seed_offset"] = so
                xs = torch.zeros(B, eng.bank.C, eng.bank.H, eng.bank.W, dtype=torch.float32, device=eng.device)
                mu = torch.empty_like(xs)
                eng.bank.strip8()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):            # warm-up: lazy bank layouts, smem attributes
                    self._run_steps(eng, steps, xs, mu, sel, sel_ls)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._run_steps(eng, steps, xs, mu, sel, sel_ls)
                while len(self._graphs) >= self.max_cached_graphs:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = (graph, xs, mu, steps, sel, sel_ls, so)
            graph, xs, mu, _steps, _sel, _sel_ls, so = self._graphs[key]
            xs.copy_(x.to(eng.device, torch.float32))
            if so is not None:
                so.copy_(torch.tensor([sto["seed"], 0], dtype=torch.int64))
            graph.replay()
            return xs.clone()

    def release_graphs(self):
        """Drops every captured trajectory graph (and the static buffers they own).  With a sharded bank the graphs hold
        captured NCCL all-gathers: release them before the process group is destroyed (distributed.shutdown)."""
        if self._graphs:
            torch.cuda.synchronize()
            self._graphs.clear()
            torch.cuda.synchronize()

    def trajectory(self, x, nsteps=None, label=None, device=None):
        """Runs the native loop eagerly and returns (final x, per-step records of x / mu) -- used by the parity
        tests, which check the denoised estimate at every step."""
        if nsteps is None:
            nsteps = self.default_time_steps if self.scales is None else len(self.scales)
        rec = []
        out = self._forward_native(x, nsteps, _as_label(label), torch.device(device or "cuda"), record=rec)
        return out, rec
