"""torch float32 CPU port of the reference's score modules -- the timed CPU baseline.

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this.  It follows the reference's operator structure
(`/root/reference/src/utils/idealscore.py`): unfold the bank batch into patches (:447), one dense contraction
q.p per batch (:454 does it as conv2d with patches as filters; here the same contraction is a matmul, which
oneDNN/MKL run at least as fast), norms (:416-418,:451), streaming softmax with a running max across batches
(:456-471), all in fp32 on the host cores, b = 1.  The reference itself cannot travel to the GPU box, so this
port is what `cpu_baseline.kind = "port"` refers to; it is pinned against the reference-generated goldens in
tests/test_oracle_golden.py::test_port_*.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def cosine_beta(t):
    return 1.0 - math.cos(t / 1.008 * math.pi / 2.0) ** 2


class _Stream:
    def __init__(self, nq, c):
        self.m = torch.full((nq,), -float("inf"))
        self.l = torch.zeros(nq)
        self.acc = torch.zeros(nq, c)

    def update(self, logits, values):                   # logits [nq, nc], values [nc, C]
        m_new = torch.maximum(self.m, logits.amax(dim=1))
        scale = torch.exp(self.m - m_new)
        p = torch.exp(logits - m_new[:, None])
        self.l = self.l * scale + p.sum(dim=1)
        self.acc = self.acc * scale[:, None] + p @ values
        self.m = m_new


def _query_patches(x, k, pad):
    d = k // 2
    xp = F.pad(x[None], (d, d, d, d), mode="circular") if pad == "circular" else F.pad(x[None], (d, d, d, d))
    return F.unfold(xp, k)[0].T.contiguous()           # [HW, D]   (:414-416 / :171-172)


def els_mu(x, bank, beta, k, logw=None, batch=64, query_pad="circular"):
    """x [C,H,W]; bank [N,C,H,W] (already label filtered); returns mu [C,H,W]."""
    c, h, w = x.shape
    a = math.sqrt(1.0 - beta)
    d = k // 2
    q = _query_patches(x, k, query_pad)
    qn = (q * q).sum(1)
    st = _Stream(h * w, c)
    for s in range(0, bank.shape[0], batch):
        imgs = bank[s:s + batch]
        p = F.unfold(imgs, k).permute(0, 2, 1).reshape(-1, q.shape[1])           # [n*P, D]  (:447-450)
        pn = (p * p).sum(1)
        vals = imgs[:, :, d:h - d, d:w - d].permute(0, 2, 3, 1).reshape(-1, c)   # centre pixels (:452)
        logits = -(qn[:, None] - 2 * a * (q @ p.T) + a * a * pn[None]) / (2 * beta)   # (:454-456)
        if logw is not None:
            npatch = p.shape[0] // imgs.shape[0]
            logits = logits + logw[s:s + batch].repeat_interleave(npatch)[None]
        st.update(logits, vals)
    return (st.acc / st.l[:, None]).T.reshape(c, h, w)


def _axis_allow(size, d):
    i = torch.arange(size)
    border = (i < d) | (i >= size - d)
    interior = (i >= d) & (i < size - d)
    return torch.where(border[:, None], i[:, None] == i[None, :], interior[None, :].expand(size, size))


def bbels_mu(x, bank, beta, k, logw=None, batch=16):
    c, h, w = x.shape
    a = math.sqrt(1.0 - beta)
    d = k // 2
    q = _query_patches(x, k, "zeros")
    qn = (q * q).sum(1)
    allow = (_axis_allow(h, d)[:, None, :, None] & _axis_allow(w, d)[None, :, None, :]).reshape(h * w, h * w)
    st = _Stream(h * w, c)
    for s in range(0, bank.shape[0], batch):
        imgs = bank[s:s + batch]
        n = imgs.shape[0]
        p = F.unfold(F.pad(imgs, (d, d, d, d)), k).permute(0, 2, 1).reshape(-1, q.shape[1])   # zero-padded patches
        pn = (p * p).sum(1)
        logits = -(qn[:, None] - 2 * a * (q @ p.T) + a * a * pn[None]) / (2 * beta)
        logits = logits.reshape(h * w, n, h * w).masked_fill(~allow[:, None, :], -float("inf")).reshape(h * w, -1)
        if logw is not None:
            logits = logits + logw[s:s + batch].repeat_interleave(h * w)[None]
        st.update(logits, imgs.permute(0, 2, 3, 1).reshape(-1, c))
    return (st.acc / st.l[:, None]).T.reshape(c, h, w)


def ls_mu(x, bank, beta, k, logw=None, batch=4096):
    c, h, w = x.shape
    a = math.sqrt(1.0 - beta)
    ones = torch.ones(1, 1, k, k)
    m = torch.full((h, w), -float("inf"))
    l = torch.zeros(h, w)
    acc = torch.zeros(c, h, w)
    for s in range(0, bank.shape[0], batch):
        imgs = bank[s:s + batch]
        e = ((x[None] - a * imgs) ** 2).sum(1, keepdim=True)                       # (:537-538)
        logits = -F.conv2d(e, ones, padding=k // 2)[:, 0] / (2 * beta)             # k x k box, zero filled (:539-541)
        if logw is not None:
            logits = logits + logw[s:s + batch][:, None, None]
        m_new = torch.maximum(m, logits.amax(0))
        scale = torch.exp(m - m_new)
        p = torch.exp(logits - m_new[None])
        l = l * scale + p.sum(0)
        acc = acc * scale[None] + (p[:, None] * imgs).sum(0)
        m = m_new
    return acc / l[None]


def mu(kind, x, bank, beta, k, logw=None, query_pad=None):
    h = x.shape[-1]
    if kind == "ELS":
        return els_mu(x, bank, beta, k, logw, query_pad=query_pad or "circular")
    if kind == "LS" or (kind == "bbELS" and k >= h):
        return ls_mu(x, bank, beta, k, logw)
    if kind == "IS":                                   # idealscore.py:560-636
        a = math.sqrt(1.0 - beta)
        logits = -((x[None] - a * bank) ** 2).sum(dim=(1, 2, 3)) / (2 * beta)
        if logw is not None:
            logits = logits + logw
        p = torch.softmax(logits, dim=0)
        return (p[:, None, None, None] * bank).sum(0)
    if kind == "bbELS":
        return bbels_mu(x, bank, beta, k, logw)
    raise ValueError(kind)


def run_machine(kind, x, bank, scales, logw=None, query_pad=None):
    """ScheduledScoreMachine.forward for one sample (idealscore.py:76-118) in the mu form."""
    nsteps = len(scales)
    x = x.clone()
    for i in range(nsteps - 1, 0, -1):
        bt = cosine_beta(i / nsteps)
        bp = max(cosine_beta(i / nsteps - 1.0 / nsteps), 0.0)
        est = mu(kind, x, bank, bt, int(scales[i]), logw, query_pad)
        r = math.sqrt(bp / bt)
        x = r * x + (math.sqrt(1 - bp) - r * math.sqrt(1 - bt)) * est
    return x
