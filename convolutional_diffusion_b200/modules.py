"""Drop-in score modules with the reference's constructor and call signatures.

Mirrors `/root/reference/src/utils/idealscore.py`:
    LocalScoreModule              :476   LS
    LocalEquivScoreModule         :375   ELS
    LocalEquivBordersScoreModule  :127   bbELS
    cosine_noise_schedule / exponential_schedule / linear_noise_schedule  :41-52
call protocol (idealscore.py:97,99; scripts/scales_calibration.py:151):
    module(t, x, label=None, device=None, k=None) -> score tensor, same shape as x
The bodies are not the reference's: the bank is uploaded once (bank.PatchBank) and each call launches the
sm_100a kernels behind include/cdscore.h.  There is no CPU path.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .bank import PatchBank, dataset_to_tensors
from .engine import ScoreEngine
from .selection import dataloader_shuffle_order


def exponential_schedule(t):
    return 1 - torch.exp(-2 * t)


def linear_noise_schedule(t):
    return 0.01 + 0.97 * t


def cosine_noise_schedule(t, mode="legacy"):
    """beta(t), the noise variance (noise_schedules.py:15-18)."""
    if mode == "legacy":
        return 1 - torch.cos(t / 1.008 * math.pi / 2) ** 2
    return 1 - torch.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2


def _as_label(label):
    if label is None:
        return None
    if torch.is_tensor(label):
        return int(label.reshape(-1)[0].item())
    return int(label)


def _label_groups(label, B):
    """Per-sample labels (an extension; the reference is b = 1 with one label per call): a label tensor
    / sequence with B > 1 entries that are not all equal -> {label: [sample indices]}; otherwise None."""
    if label is None or B <= 1:
        return None
    flat = label.reshape(-1).tolist() if torch.is_tensor(label) else (list(label) if hasattr(label, "__len__") else None)
    if flat is None or len(flat) == 1:
        return None
    if len(flat) != B:
        raise ValueError(f"label has {len(flat)} entries for a batch of {B} samples (expected 1 or {B})")
    groups = {}
    for i, v in enumerate(flat):
        groups.setdefault(int(v), []).append(i)
    return groups if len(groups) > 1 else None


class _DifferentiableScore(torch.autograd.Function):
    """score(x) with a closed-form backward (cds_score_vjp_simt), so that callers which differentiate the score modules
    with autograd -- the exterior-derivative analysis, src/utils/exterior_derivative.py:68-79 and
    scripts/analyze_exterior_derivative.py:171-183 -- work on the CUDA path.  Forward and backward use the exact-fp32 SIMT
    kernels (one consistent set of softmax statistics); not differentiable twice."""

    @staticmethod
    def forward(ctx, x, eng, kind, pad, beta, k, sel):
        from . import _lib
        with torch.cuda.device(eng.device):
            P = eng.combine(eng.simt_partials(kind, pad, x, beta, k, sel, tag="grad"))
            mu, score = torch.empty_like(x), torch.empty_like(x)
            eng.finalize(P, x, beta, mu, score)
            ctx.save_for_backward(x, beta, P.m[0].clone(), P.l[0].clone(), mu)
        ctx.args = (eng, kind, pad, k, sel)
        return score

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        from . import _lib
        x, beta, m, l, mu = ctx.saved_tensors
        eng, kind, pad, k, sel = ctx.args
        idx, logw, n_sel = sel
        b = eng.bank
        g = g.contiguous().float()
        grad_mu = torch.zeros_like(x)
        with torch.cuda.device(eng.device):
            if n_sel > 0:
                S = eng._splits((b.H * b.W + 127) // 128, x.shape[0], n_sel, waves=4, min_images=8)
                _lib.check(eng.lib.cds_score_vjp_simt(_lib.KIND[kind], _lib.PAD[pad], _lib.ptr(x), x.shape[0], b.C, b.H, b.W, k,
                                                      _lib.ptr(beta), _lib.ptr(b.images), _lib.ptr(idx), _lib.ptr(logw), n_sel, S,
                                                      _lib.ptr(m), _lib.ptr(l), _lib.ptr(mu), _lib.ptr(g), _lib.ptr(grad_mu),
                                                      _lib.stream_ptr()), "cds_score_vjp_simt")
                eng.launches += 1
            if eng.group is not None:                      # every rank holds the contributions of its own bank slice
                import torch.distributed as dist
                dist.all_reduce(grad_mu, group=eng.group)
        bb = beta.view(-1, 1, 1, 1)
        return -g / bb + (torch.sqrt(1 - bb) / bb) * grad_mu, None, None, None, None, None, None


class _ScoreModuleBase(nn.Module):
    kind = None
    query_pad = None

    def __init__(self, dataset, kernel_size, batch_size, image_size, schedule, max_samples, shuffle,
                 precision="f16x2", use_tensor_cores=True, process_group=None, bank=None):
        super().__init__()
        self.dataset = dataset
        self.batch_size = batch_size
        self.kernel_size = kernel_size
        self.image_size = image_size
        self.schedule = schedule
        self.max_samples = max_samples
        self.shuffle = shuffle
        self.precision = precision
        self.use_tensor_cores = use_tensor_cores
        self.process_group = process_group
        self._bank = bank
        self._engine = None

    # -- lazily built device state ----------------------------------------------------------------
    def engine(self, device=None) -> ScoreEngine:
        if self._engine is not None and device is not None:
            want = torch.device(device)
            if want.type == "cuda" and want.index is not None and want != self._engine.device:
                raise RuntimeError(f"this score module's bank lives on {self._engine.device}; it cannot be evaluated on "
                                   f"{want} (build a second module, or pass bank= a PatchBank of that device)")
        if self._engine is None:
            if self._bank is None:
                images, labels = dataset_to_tensors(self.dataset)
                rank, world = self._rank_world()
                self._bank = PatchBank(images, labels, device=device, rank=rank, world=world)
            elif (self._bank.rank, self._bank.world) != self._rank_world():
                raise RuntimeError("bank= was built for another sharding (rank/world) than this module's process_group")
            self._engine = ScoreEngine(self._bank, precision=self.precision,
                                       use_tensor_cores=self.use_tensor_cores, group=self.process_group)
        return self._engine

    @property
    def bank(self):
        return self.engine().bank

    def _rank_world(self):
        if self.process_group is None:
            return 0, 1
        import torch.distributed as dist
        return dist.get_rank(self.process_group), dist.get_world_size(self.process_group)

    def selection(self, label, kind=None):
        kind = kind or self.kind
        eng = self.engine()
        order = None
        rank, world = self._rank_world()
        if self.shuffle or kind == "LS" and self._ls_shuffles():
            order = dataloader_shuffle_order(eng.bank.N)
            if world > 1:
                # every rank must split the SAME visiting order: ranks whose global RNG states differ would otherwise count
                # some images twice and drop others.  Rank 0's draw wins (distributed.shared_order).
                from .distributed import shared_order
                order = shared_order(order, self.process_group, eng.device)
        return eng.bank.selection(kind, label, self.batch_size, self.max_samples, order)

    def _ls_shuffles(self):
        """LS hard-codes shuffle=True (idealscore.py:489); the order only matters when batches end up with
        unequal post-filter sizes, so skip the permutation when a single batch covers the bank."""
        return self.batch_size < self.engine().bank.N          # not len(dataset): an (images, labels) pair has length 2

    def betas(self, t, device):
        return self.schedule(torch.as_tensor(t, dtype=torch.float32).reshape(-1).cpu()).to(device, torch.float32)

    def forward(self, t, x, label=None, device=None, k=None):
        if device is None:
            device = torch.device("cuda")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the B200 score modules run on CUDA only (no CPU fallback)")
        eng = self.engine(device)
        k = self.kernel_size if k is None else int(k)
        xd = x.to(eng.device, torch.float32).contiguous()
        tt = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        if tt.numel() == 1 and xd.shape[0] > 1:
            tt = tt.expand(xd.shape[0])
        groups = _label_groups(label, xd.shape[0])
        if groups is not None:                   # one evaluation per distinct label, results scattered back
            out = torch.empty_like(xd)
            for lab_g, rows in groups.items():
                sel_rows = torch.as_tensor(rows, device=xd.device)
                out[sel_rows] = self.forward(tt[rows], xd[sel_rows], label=lab_g, device=device, k=k)
            return out
        beta_cpu = self.schedule(tt.cpu().float())
        beta = beta_cpu.to(eng.device, torch.float32).contiguous()
        lab = _as_label(label)
        sel = self.selection(lab)
        sel_ls = None
        if self.kind == "bbELS" and k >= xd.shape[-2]:
            sel_ls = self.selection(lab, kind="LS")
        if xd.requires_grad and torch.is_grad_enabled():
            # autograd through the module (exterior-derivative callers): exact-fp32 forward + closed-form backward
            kind, kk, pad = self.kind, k, (self.query_pad or "zeros")
            if kind == "IS":
                kind, kk, pad = "LS", 2 * max(eng.bank.H, eng.bank.W) - 1, "zeros"
            elif kind == "bbELS" and k >= eng.bank.H:
                kind, sel, pad = "LS", sel_ls, "zeros"
            elif kind == "LS":
                pad = "zeros"
            return _DifferentiableScore.apply(xd, eng, kind, pad, beta, kk, sel)
        score = torch.empty_like(xd)
        eng.evaluate(self.kind, xd, beta, k, sel, query_pad=self.query_pad, mu=None, score=score,
                     beta_min=float(beta_cpu.min()), sel_ls=sel_ls)
        return score


    def forward_multi_k(self, t, x, kernelsizes, label=None, device=None):
        """The scores of the SAME (t, x) for several kernel sizes, [len(kernelsizes), B, C, H, W]: what one step of the
        kernel-size calibration needs (scripts/scales_calibration.py:151 calls one module per size).  x is uploaded, the
        noise levels are formed and the bank selection (label filter, log-weights) is resolved once for all sizes; the
        per-size launches follow each other on the stream with no host synchronisation in between, and B samples share
        every pass over the bank."""
        if device is None:
            device = torch.device("cuda")
        eng = self.engine(torch.device(device))
        xd = x.to(eng.device, torch.float32).contiguous()
        tt = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        if tt.numel() == 1 and xd.shape[0] > 1:
            tt = tt.expand(xd.shape[0])
        groups = _label_groups(label, xd.shape[0])
        out = torch.empty((len(kernelsizes),) + tuple(xd.shape), dtype=torch.float32, device=eng.device)
        if groups is not None:                   # per-sample labels: one pass per distinct label
            for lab_g, rows in groups.items():
                sel_rows = torch.as_tensor(rows, device=xd.device)
                out[:, sel_rows] = self.forward_multi_k(tt[rows], xd[sel_rows], kernelsizes, label=lab_g, device=device)
            return out
        beta_cpu = self.schedule(tt.cpu().float())
        beta = beta_cpu.to(eng.device, torch.float32).contiguous()
        beta_min = float(beta_cpu.min())
        lab = _as_label(label)
        sel = self.selection(lab)
        for q, k in enumerate(kernelsizes):
            k = int(k)
            sel_ls = self.selection(lab, kind="LS") if (self.kind == "bbELS" and k >= xd.shape[-2]) else None
            eng.evaluate(self.kind, xd, beta, k, sel, query_pad=self.query_pad, mu=None, score=out[q], beta_min=beta_min,
                         sel_ls=sel_ls)
        return out


class LocalScoreModule(_ScoreModuleBase):
    """LS (idealscore.py:476-557).  Note the reference default schedule is the exponential one (:484)."""
    kind = "LS"

    def __init__(self, dataset, kernel_size=3, image_size=32, batch_size=256, show_plots=False,
                 schedule=exponential_schedule, max_samples=None, **kwargs):
        super().__init__(dataset, kernel_size, batch_size, image_size, schedule, max_samples, shuffle=False, **_own(kwargs))
        self.show_plots = show_plots


class LocalEquivScoreModule(_ScoreModuleBase):
    """ELS (idealscore.py:375-473): circular-padded query, every valid patch of every image."""
    kind = "ELS"
    query_pad = "circular"

    def __init__(self, dataset, kernel_size=3, batch_size=64, image_size=32, channels=3,
                 schedule=cosine_noise_schedule, max_samples=None, shuffle=False, query_pad="circular", **kwargs):
        super().__init__(dataset, kernel_size, batch_size, image_size, schedule, max_samples, shuffle, **_own(kwargs))
        self.channels = channels
        self.query_pad = query_pad       # "zeros" = the zero-padded ELS variant of BASELINE config 2


class LocalEquivBordersScoreModule(_ScoreModuleBase):
    """bbELS (idealscore.py:127-372): zero-padded query, border-aware candidate sets; k >= H -> LS (:163)."""
    kind = "bbELS"
    query_pad = "zeros"

    def __init__(self, dataset, kernel_size=3, batch_size=64, image_size=32, channels=3,
                 schedule=cosine_noise_schedule, max_samples=None, shuffle=False, **kwargs):
        super().__init__(dataset, kernel_size, batch_size, image_size, schedule, max_samples, shuffle, **_own(kwargs))
        self.channels = channels


class IdealScoreModule(_ScoreModuleBase):
    """IS (idealscore.py:560-636): whole-image posterior mean.  It equals LS with a window that covers the whole
    image from every pixel (k = 2*max(H,W) - 1), so it runs on the LS bank-streaming kernel; `k` is ignored as in the
    reference (`forward(t, x, label=None, device=None, **kwargs)`)."""
    kind = "IS"

    def __init__(self, dataset, image_size=32, batch_size=128, schedule=cosine_noise_schedule, max_samples=None,
                 shuffle=False, **kwargs):
        super().__init__(dataset, None, batch_size, image_size, schedule, max_samples, shuffle, **_own(kwargs))

    def forward(self, t, x, label=None, device=None, **kwargs):
        k = 2 * max(int(x.shape[-1]), int(x.shape[-2])) - 1
        return super().forward(t, x, label=label, device=device, k=k)


def _own(kwargs):
    return {k: kwargs[k] for k in ("precision", "use_tensor_cores", "process_group", "bank") if k in kwargs}
