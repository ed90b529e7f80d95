"""float64 numpy restatement of the reference's analytic score modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates; `/root/reference/` is implied in all
citations.  The three modules are written in the *unified masked-softmax
form* (SURVEY.md §8 N3):

    mu(i,j) = softmax_{(n,u,v) in S(i,j), y_n == label}
                 [ -||q(i,j) - a p(n,u,v)||^2 / (2 beta) + logw_n ] . T_n[:, u, v]
    score   = -(x - a mu) / beta

with  a = sqrt(1-beta),  q/p = k x k x C patches centred at the pixel:
  LS    S = {(u,v) == (i,j)}, q and p zero padded            (idealscore.py:497-557)
  ELS   S = {interior (u,v)}, q circular padded, p un-padded (idealscore.py:397-473)
  bbELS S = axis-wise border rule, q and p zero padded       (idealscore.py:156-372)

All arithmetic here is float64 and brute force: it is a checker, not a
fast path.
"""
from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

__all__ = [
    "cosine_beta", "exponential_beta", "select_bank", "ls_mu", "els_mu", "bbels_mu",
    "score_from_mu", "score", "machine_coeffs", "run_machine", "pair_count",
]


# --------------------------------------------------------------------------
# noise schedules (src/utils/noise_schedules.py:5-18, idealscore.py:41-52)
# --------------------------------------------------------------------------
def cosine_beta(t, mode="legacy"):
    """beta(t): noise *variance*.  noise_schedules.py:15-18 ('legacy' is the default)."""
    t = np.asarray(t, dtype=np.float64)
    if mode == "legacy":
        return 1.0 - np.cos(t / 1.008 * math.pi / 2.0) ** 2
    return 1.0 - np.cos((t + 0.008) / 1.008 * math.pi / 2.0) ** 2


def exponential_beta(t):
    """noise_schedules.py:5-9 / idealscore.py:41-42 (LS's constructor default)."""
    return 1.0 - np.exp(-2.0 * np.asarray(t, dtype=np.float64))


# --------------------------------------------------------------------------
# which bank images take part, and with which weight
# --------------------------------------------------------------------------
def select_bank(kind, labels, label, batch_size, max_samples, order=None):
    """Replays the reference's DataLoader loop bookkeeping.

    Returns (idx, logw): indices into the dataset of the images that
    contribute, in visiting order, and the per-image log-weight.

    ELS   idealscore.py:430-444,470-471  pre-filter running count, `break`
          once it exceeds max_samples; per-batch torch.mean => weight 1/n_b.
    LS    idealscore.py:521-535,553-554  post-filter running count; mean quirk.
    bbELS idealscore.py:184-193,336-370  counter advances by batch_size per
          visited batch; plain torch.sum => weight 1.
    `order` is the DataLoader's visiting order (identity when shuffle=False).
    """
    labels = np.asarray(labels)
    n = len(labels)
    order = np.arange(n) if order is None else np.asarray(order)
    idx, logw = [], []
    seen = 0
    for s in range(0, n, batch_size):
        b = order[s:s + batch_size]
        if kind == "ELS":
            seen += len(b)
            if max_samples is not None and seen > max_samples:
                break
            if label is not None:
                b = b[labels[b] == label]
            if len(b) == 0:
                continue
            w = -math.log(len(b))
        elif kind in ("LS", "IS"):
            if label is not None:
                b = b[labels[b] == label]
            if len(b) == 0:
                continue
            seen += len(b)
            if max_samples is not None and seen > max_samples:
                break
            w = -math.log(len(b))
        elif kind == "bbELS":
            if max_samples is not None and seen > max_samples:
                break
            seen += batch_size
            if label is not None:
                b = b[labels[b] == label]
            if len(b) == 0:
                continue
            w = 0.0
        else:
            raise ValueError(kind)
        idx.extend(int(i) for i in b)
        logw.extend([w] * len(b))
    return np.asarray(idx, dtype=np.int64), np.asarray(logw, dtype=np.float64)


# --------------------------------------------------------------------------
# patch helpers
# --------------------------------------------------------------------------
def _pad(img, d, mode):
    """img [C,H,W] -> [C,H+2d,W+2d]; 'circular' = F.pad(mode='circular') (idealscore.py:35,414),
    'zeros' = F.pad(value=0) (idealscore.py:171)."""
    if d == 0:
        return img
    if mode == "circular":
        return np.pad(img, ((0, 0), (d, d), (d, d)), mode="wrap")
    return np.pad(img, ((0, 0), (d, d), (d, d)), mode="constant")


def _windows(img, k):
    """All k x k windows of img [C,Hp,Wp] as rows: [Hp-k+1, Wp-k+1, C*k*k] (F.unfold order c,dy,dx;
    idealscore.py:172,416,447 -- the order is irrelevant to the result as long as q and p agree)."""
    w = sliding_window_view(img, (k, k), axis=(1, 2))          # [C,H',W',k,k]
    hp, wp = w.shape[1], w.shape[2]
    return np.ascontiguousarray(w.transpose(1, 2, 0, 3, 4)).reshape(hp, wp, -1)


class _Softmax:
    """Streaming (max, sum-exp, weighted-sum) accumulator in float64; the
    mathematical content of idealscore.py:458-471 without its batching."""

    def __init__(self, nq, c):
        self.m = np.full(nq, -np.inf)
        self.l = np.zeros(nq)
        self.acc = np.zeros((nq, c))

    def update(self, logits, values):
        """logits [nq, nc] (may contain -inf), values [nc, C]."""
        cm = logits.max(axis=1)
        m_new = np.maximum(self.m, cm)
        safe = np.where(np.isfinite(m_new), m_new, 0.0)
        scale = np.where(np.isfinite(self.m), np.exp(self.m - safe), 0.0)
        p = np.exp(logits - safe[:, None])
        self.l = self.l * scale + p.sum(axis=1)
        self.acc = self.acc * scale[:, None] + p @ values
        self.m = m_new

    def mean(self):
        return self.acc / self.l[:, None]


def _prep(x, bank, beta, logw):
    x = np.asarray(x, dtype=np.float64)
    bank = np.asarray(bank, dtype=np.float64)
    assert x.ndim == 3 and bank.ndim == 4 and bank.shape[1:] == x.shape, (x.shape, bank.shape)
    beta = float(beta)
    a = math.sqrt(1.0 - beta)
    logw = np.zeros(len(bank)) if logw is None else np.asarray(logw, dtype=np.float64)
    return x, bank, beta, a, logw


# --------------------------------------------------------------------------
# the three score machines: denoised estimate mu [C,H,W]
# --------------------------------------------------------------------------
def els_mu(x, bank, beta, k, logw=None, query_pad="circular", chunk=16):
    """ELS (idealscore.py:397-473).  x [C,H,W]; bank [N,C,H,W] already label filtered.
    Query patches come from the circular-padded x (:414-416); candidates are all valid
    (un-padded) k x k patches of every bank image (:447-450) with value = centre pixel (:452)."""
    x, bank, beta, a, logw = _prep(x, bank, beta, logw)
    c, h, w = x.shape
    d = k // 2
    q = _windows(_pad(x, d, query_pad), k).reshape(h * w, -1)          # [HW, D]
    qn = (q * q).sum(1)
    sm = _Softmax(h * w, c)
    for s in range(0, len(bank), chunk):
        ps, cs, ws = [], [], []
        for n in range(s, min(s + chunk, len(bank))):
            p = _windows(bank[n], k)                                    # [Ph,Pw,D]
            ps.append(p.reshape(-1, p.shape[-1]))
            cs.append(bank[n][:, d:h - d, d:w - d].reshape(c, -1).T)   # centre pixels [P,C]
            ws.append(np.full(ps[-1].shape[0], logw[n]))
        p = np.concatenate(ps); cv = np.concatenate(cs); lw = np.concatenate(ws)
        pn = (p * p).sum(1)
        logits = -(qn[:, None] - 2.0 * a * (q @ p.T) + a * a * pn[None, :]) / (2.0 * beta) + lw[None, :]
        sm.update(logits, cv)
    return sm.mean().T.reshape(c, h, w)


def _axis_rule(i, u, size, d):
    """bbELS candidate rule along one axis (idealscore.py:206-288): a query whose
    patch crosses the border (i<d or i>=size-d) only sees candidates at the same coordinate;
    interior queries see every interior coordinate."""
    border = (i < d) | (i >= size - d)
    interior_u = (u >= d) & (u < size - d)
    return np.where(border[:, None], i[:, None] == u[None, :], interior_u[None, :])


def bbels_mu(x, bank, beta, k, logw=None, chunk=8):
    """bbELS (idealscore.py:156-372).  Zero-padded query (:171-174); candidate = zero-padded patch
    of a bank image centred at (u,v) subject to the axis-wise rule; plain sum weights (:336-368).
    Callers handle the k >= H delegation to LS (:163-164)."""
    x, bank, beta, a, logw = _prep(x, bank, beta, logw)
    c, h, w = x.shape
    d = k // 2
    assert k < h and k < w
    q = _windows(_pad(x, d, "zeros"), k).reshape(h * w, -1)
    qn = (q * q).sum(1)
    ii, jj = np.divmod(np.arange(h * w), w)
    # [HW queries, HW candidate centres]
    allow = _axis_rule(ii, np.arange(h), h, d)[:, :, None] & _axis_rule(jj, np.arange(w), w, d)[:, None, :]
    allow = allow.reshape(h * w, h * w)
    sm = _Softmax(h * w, c)
    for s in range(0, len(bank), chunk):
        logits, vals = [], []
        for n in range(s, min(s + chunk, len(bank))):
            p = _windows(_pad(bank[n], d, "zeros"), k).reshape(h * w, -1)
            pn = (p * p).sum(1)
            lg = -(qn[:, None] - 2.0 * a * (q @ p.T) + a * a * pn[None, :]) / (2.0 * beta) + logw[n]
            logits.append(np.where(allow, lg, -np.inf))
            vals.append(bank[n].reshape(c, -1).T)
        sm.update(np.concatenate(logits, axis=1), np.concatenate(vals))
    return sm.mean().T.reshape(c, h, w)


def ls_mu(x, bank, beta, k, logw=None):
    """LS (idealscore.py:497-557).  Same-location candidates only; the k x k window of squared
    differences is zero-filled outside the image (`unfold(..., padding=k//2)` :539), i.e. both
    patches are zero padded."""
    x, bank, beta, a, logw = _prep(x, bank, beta, logw)
    c, h, w = x.shape
    d = k // 2
    e = ((x[None] - a * bank) ** 2).sum(1)                              # [N,H,W]  (:537-538)
    e = np.pad(e, ((0, 0), (d, d), (d, d)))
    box = sliding_window_view(e, (k, k), axis=(1, 2)).sum(axis=(-1, -2))  # [N,H,W]  (:539-541)
    logits = -box / (2.0 * beta) + logw[:, None, None]
    m = logits.max(0)
    p = np.exp(logits - m[None])
    return (p[:, None] * bank).sum(0) / p.sum(0)[None]


def is_mu(x, bank, beta, logw=None):
    """IS (idealscore.py:560-636): whole-image posterior mean, softmax over images of -||x - a T_n||^2 / (2 beta).
    Identical to LS with a window that covers the whole image from every pixel (k >= 2H-1)."""
    x, bank, beta, a, logw = _prep(x, bank, beta, logw)
    dist = ((x[None] - a * bank) ** 2).sum(axis=(1, 2, 3))
    logits = -dist / (2.0 * beta) + logw
    p = np.exp(logits - logits.max())
    return np.tensordot(p, bank, axes=1) / p.sum()


def score_from_mu(x, mu, beta):
    """score = -(x - a mu)/beta  (idealscore.py:372,473,557)."""
    beta = float(beta)
    return -(np.asarray(x, np.float64) - math.sqrt(1.0 - beta) * mu) / beta


def score(kind, x, bank, beta, k, logw=None):
    """Score of one sample; bbELS delegates to LS when k >= H (idealscore.py:163-164)."""
    h = x.shape[-1]
    if kind == "ELS":
        mu = els_mu(x, bank, beta, k, logw)
    elif kind == "bbELS":
        mu = ls_mu(x, bank, beta, k, logw) if k >= h else bbels_mu(x, bank, beta, k, logw)
    elif kind == "LS":
        mu = ls_mu(x, bank, beta, k, logw)
    elif kind == "IS":
        mu = is_mu(x, bank, beta, logw)
    else:
        raise ValueError(kind)
    return score_from_mu(x, mu, beta), mu


# --------------------------------------------------------------------------
# the sampler (idealscore.py:76-118)
# --------------------------------------------------------------------------
def machine_coeffs(nsteps, schedule=cosine_beta):
    """Per-step (i, beta_t, beta_prev, c_x, c_eps) of the deterministic DDIM update
    x <- c_x * x + c_eps * eps,  eps = -sqrt(beta_t) * score   (idealscore.py:88-116)."""
    out = []
    for i in range(nsteps - 1, 0, -1):
        bt = float(schedule(i / nsteps))
        bp = float(schedule(i / nsteps - 1.0 / nsteps))
        bp = max(bp, 0.0)
        ratio = math.sqrt((1.0 - bp) / (1.0 - bt))
        out.append((i, bt, bp, ratio, math.sqrt(bp) - ratio * math.sqrt(bt)))
    return out


def run_machine(kind, x, bank, scales, logw=None, nsteps=None, default_k=3, schedule=cosine_beta,
                record=None):
    """ScheduledScoreMachine.forward for one sample x [C,H,W] (idealscore.py:76-118).
    `scales[i]` is the kernel size used at t=i/nsteps (:95); scales[0] is never read."""
    x = np.array(x, dtype=np.float64)
    if nsteps is None:
        nsteps = len(scales) if scales is not None else 20
    for i, bt, _bp, cx, ce in machine_coeffs(nsteps, schedule):
        k = default_k if scales is None else int(scales[i])
        s, mu = score(kind, x, bank, bt, k, logw)
        if record is not None:
            record.append(dict(i=i, k=k, beta=bt, x=x.copy(), score=s, mu=mu))
        eps = -math.sqrt(bt) * s
        x = cx * x + ce * eps
    return x


def ddpm_coeffs(nsteps, schedule=cosine_beta, first_step=None):
    """Per-step (i, beta_t, c_mu, c_eps, sigma) of the stochastic branch of the reference sampler (src/models.py:48-64):
        sigma = sqrt(b'/b) sqrt(1 - a/a'),  x <- sqrt(a') * (x - sqrt(b) eps)/sqrt(a) + sqrt(1 - a' - sigma^2) eps + sigma z
    The loop of DDIM.sample starts at i = nsteps (first_step = None), ScheduledScoreMachine's at nsteps - 1."""
    out = []
    for i in range(nsteps if first_step is None else first_step, 0, -1):
        bt = float(schedule(i / nsteps))
        bp = max(float(schedule(i / nsteps - 1.0 / nsteps)), 0.0)
        at, ap = 1.0 - bt, 1.0 - bp
        sigma = math.sqrt(bp / bt) * math.sqrt(max(1.0 - at / ap, 0.0))
        out.append((i, bt, math.sqrt(ap), math.sqrt(max(1.0 - ap - sigma * sigma, 0.0)), sigma))
    return out


def run_machine_ddpm(kind, x, bank, k, nsteps, noises, logw=None, schedule=cosine_beta, first_step=None, scales=None):
    """DDIM.sample(ddpm=True) (src/models.py:48-64) for one sample with eps = -sqrt(beta) * score of the analytic module;
    noises[q] [C,H,W] is the Gaussian draw of the q-th step."""
    x = np.array(x, dtype=np.float64)
    for q, (i, bt, c_mu, c_eps, sigma) in enumerate(ddpm_coeffs(nsteps, schedule, first_step)):
        kk = k if scales is None else int(scales[min(i, len(scales) - 1)])
        s, mu = score(kind, x, bank, bt, kk, logw)
        eps = -math.sqrt(bt) * s
        x = c_mu * (x - math.sqrt(bt) * eps) / math.sqrt(1.0 - bt) + c_eps * eps + sigma * np.asarray(noises[q], dtype=np.float64)
    return x


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 numpy arrays: counter [..., 4], key [..., 2] -> [..., 4].  Restates the
    generator of csrc/simt_kernels.cu so that the device noise stream of the stochastic sampler can be checked bit for bit."""
    c = [np.asarray(counter[..., j], dtype=np.uint64) for j in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & m32
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & m32
        c = [n0, p1 & m32, n2, p0 & m32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal(n, seed, ctr):
    """The n standard normals cds_randn_philox / cds_finish draw for elements 0..n-1 at stream position ctr."""
    e = np.arange(n, dtype=np.uint64)
    counter = np.stack([e & np.uint64(0xFFFFFFFF), e >> np.uint64(32), np.full(n, ctr & 0xFFFFFFFF, np.uint64),
                        np.full(n, ctr >> 32, np.uint64)], axis=-1)
    key = np.stack([np.full(n, seed & 0xFFFFFFFF, np.uint64), np.full(n, seed >> 32, np.uint64)], axis=-1)
    w = philox4x32_10(counter, key).astype(np.float64)
    u1 = np.minimum((w[:, 0] + 0.5) * 2.0 ** -32, 1.0)
    u2 = (w[:, 1] + 0.5) * 2.0 ** -32
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def pair_count(kind, h, w, k, n):
    """Number of (query pixel, candidate patch) pairs of one score evaluation."""
    d = k // 2
    if kind == "LS" or (kind == "bbELS" and k >= h):
        return h * w * n
    if kind == "ELS":
        return h * w * n * (h - k + 1) * (w - k + 1)
    ih, iw = h - 2 * d, w - 2 * d
    return n * (ih * iw * ih * iw + 2 * d * iw * iw + 2 * d * ih * ih + 4 * d * d)
