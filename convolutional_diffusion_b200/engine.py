"""Launch layer between the Python score modules and the C ABI: picks the kernel for a (kind, geometry),
sizes the grid against the SM count, owns the partial buffers and runs merge + epilogue.

Everything here is stream ordered on torch's current stream and free of host synchronisation, so a whole
trajectory can be captured in one CUDA graph (machine.py).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from .bank import PatchBank

_SM = {}


def sm_count(device):
    key = torch.device(device).index or 0
    if key not in _SM:
        _SM[key] = torch.cuda.get_device_properties(device).multi_processor_count
    return _SM[key]


class Partials:
    """(m, l, acc) for S slices of B samples (layout in include/cdscore.h)."""

    def __init__(self, S, B, C, HW, device):
        self.S, self.B, self.C, self.HW = S, B, C, HW
        self.m = torch.zeros(S, B, HW, dtype=torch.float32, device=device)
        self.l = torch.zeros(S, B, HW, dtype=torch.float32, device=device)
        self.acc = torch.zeros(S, B, C, HW, dtype=torch.float32, device=device)

    def packed(self):
        return torch.cat([self.m[0].reshape(self.B, 1, self.HW), self.l[0].reshape(self.B, 1, self.HW),
                          self.acc[0]], dim=1)


class ScoreEngine:
    """Evaluates one score module kind against a PatchBank."""

    def __init__(self, bank: PatchBank, precision="f16x2", use_tensor_cores=True, group=None):
        self.bank = bank
        self.lib = bank.lib
        self.device = bank.device
        self.precision = precision
        self.use_tensor_cores = use_tensor_cores
        self.group = group                       # torch.distributed process group for bank sharding (or None)
        # epilogue of the ELS tensor-core kernel: "auto" (P.V contraction on the tensor cores where supported and measured
        # faster: k <= 9 and a/beta <= 100), "fma" (weighted sum on the FMA pipe), "pv" (P.V wherever supported); CDS_ELS_VARIANT overrides
        self.els_variant = os.environ.get("CDS_ELS_VARIANT", "auto")
        self.centre_window = os.environ.get("CDS_CENTRE_WINDOW", "1") != "0"   # A/B switch: 0 = centre kernel over all pixels
        self.ls_variant = os.environ.get("CDS_LS_VARIANT", "auto")           # "simt": keep LS off the tensor cores
        self.edge_variant = os.environ.get("CDS_EDGE_VARIANT", "auto")       # "simt": keep the bbELS edge bands off the tensor cores
        if self.els_variant not in _lib.ELS_VARIANT:
            raise ValueError(f"CDS_ELS_VARIANT must be one of {sorted(_lib.ELS_VARIANT)}, got {self.els_variant!r}")
        self._buf = {}
        self.launches = 0                        # kernels launched through this engine (bench bookkeeping)

    # ---- helpers -------------------------------------------------------------------------------
    def _partials(self, tag, S, B):
        key = (tag, S, B)
        if key not in self._buf:
            self._buf[key] = Partials(S, B, self.bank.C, self.bank.H * self.bank.W, self.device)
        return self._buf[key]

    def passes_for(self, k, beta_min):
        """Tensor-core passes over the query: 1 = fp16 query (11-bit significand), 2 = fp16 hi + lo residual
        (22 bits, fp32-grade).  The query rounding error of one pass enters the logits as ~2^-12 (a/beta) sqrt(D),
        D = k*k*C, so "auto" uses one pass while (a/beta) * sqrt(D) <= 20 (tightened to 20 * 30/sqrt(D) for D > 900).
        Measured against the float64 oracle (tolerance 1e-3 on mu): along the headline schedule
        (tests/gpu_step_errors.py, 60-image bank, 4 seeds; profiles/r01g_step_errors.log) a single pass costs
        <= 2.9e-4 inside that range, 2.0-4.1e-4 up to 30, 7.8e-4 at 38 (k=7, a/beta = 3.2) and 1.5e-3 at 850 (k=3,
        a/beta = 165); on a 21-image bank with k=27 (tests/gpu_case_probe.py) 1.4e-3 at 27 and 2.0e-4 at 8.  Outside
        the range the second pass brings the error to ~1e-5."""
        if self.precision == "f16":
            return 1
        if self.precision == "f16x2":
            return 2
        a_over_b = (max(1.0 - beta_min, 0.0) ** 0.5) / max(beta_min, 1e-6)
        root_d = k * self.bank.C ** 0.5
        return 1 if a_over_b * root_d <= 20.0 * min(1.0, 30.0 / root_d) else 2

    @staticmethod
    def _a_over_beta(beta_min):
        return (max(1.0 - beta_min, 0.0) ** 0.5) / max(beta_min, 1e-6)

    def umma_supported(self, k, passes):
        b = self.bank
        if not self.use_tensor_cores:
            return False
        planes = 1 if b.strip8()[1] is None else 2
        return self.lib.cds_els_umma_smem_bytes(b.C, b.H, b.W, k, passes, planes) > 0

    def _splits(self, tiles, B, n_sel, waves=4, min_images=16):
        """Bank slices per (query tile, sample) so that the grid fills whole waves of SMs: picks the split count
        with the best wave efficiency ctas / (ceil(ctas/SMs)*SMs) among up to `waves` full waves, preferring fewer CTAs
        on ties, and never leaves a CTA with fewer than min_images images (fixed per-CTA setup cost)."""
        sms = sm_count(self.device)
        base = max(1, tiles * B)
        best, best_eff = 1, 0.0
        smax = max(1, min(n_sel // max(1, min_images), (sms * waves) // base))
        for s in range(1, smax + 1):
            ctas = base * s
            eff = ctas / (((ctas + sms - 1) // sms) * sms)
            if eff > best_eff + 0.03:
                best, best_eff = s, eff
        return int(max(1, min(best, n_sel)))

    def _empty_shard(self, tag, B):
        """A rank whose interleaved shard holds no image of the requested class contributes the neutral element of the
        log-sum-exp merge (m = -inf, l = 0, acc = 0); the kernels are not launched."""
        P = self._partials(tag, 1, B)
        P.m.fill_(float("-inf"))
        P.l.zero_()
        P.acc.zero_()
        return P

    # ---- kernels -------------------------------------------------------------------------------
    def simt_partials(self, kind, pad, x, beta, k, sel, region=0, tag="simt"):
        idx, logw, n_sel = sel
        if n_sel == 0:
            return self._empty_shard(tag, x.shape[0])
        b = self.bank
        B = x.shape[0]
        tiles = (b.H * b.W + 127) // 128
        S = self._splits(tiles, B, n_sel, waves=4, min_images=8)
        P = self._partials(tag, S, B)
        _lib.check(self.lib.cds_partials_simt(_lib.KIND[kind], _lib.PAD[pad], _lib.ptr(x), B, b.C, b.H, b.W, k,
                                              _lib.ptr(beta), _lib.ptr(b.images), _lib.ptr(idx), _lib.ptr(logw),
                                              n_sel, S, region, _lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc),
                                              _lib.stream_ptr()), "cds_partials_simt")
        self.launches += 1
        return P

    def simt_supported(self, k):
        """The generic kernel keeps the padded query and one padded bank image in shared memory."""
        b = self.bank
        d = k // 2
        return 2 * b.C * (b.H + 2 * d) * (b.W + 2 * d) * 4 <= 227 * 1024

    def ls_supported(self, k):
        b = self.bank
        if self.lib.cds_ls_rows_supported(b.C, b.H, b.W, k):
            return True
        d = k // 2
        if d >= max(b.H, b.W) - 1:                 # whole-image window (IS): block-wide sums, no box filter
            return b.C in (1, 3) and b.H * b.W <= 4096
        smem = (8 if b.C == 1 else 4) * 4 * (b.H * (b.W + 2 * d) + (b.H + 2 * d) * b.W)   # images per round
        return b.C in (1, 3) and b.H * b.W <= 4096 and smem <= 227 * 1024

    def ls_umma_supported(self, k, passes):
        """Tensor-core LS: single-channel 8-bit-exact bank, a window smaller than the image, query band fits shared memory."""
        b = self.bank
        return (self.use_tensor_cores and self.ls_variant != "simt" and b.C == 1 and b.strip8()[1] is None
                and self.lib.cds_ls_umma_smem_bytes(b.C, b.H, b.W, k, passes) > 0)

    def ls_partials(self, x, beta, k, sel, tag="ls", passes=None):
        """LS partials: the tensor-core kernel (csrc/ls_umma.cu) when `passes` is given and the geometry is supported, else the
        bank-streaming SIMT kernels (csrc/ls_rows_kernel.cu, csrc/ls_kernel.cu)."""
        idx, logw, n_sel = sel
        if n_sel == 0:
            return self._empty_shard(tag, x.shape[0])
        b = self.bank
        B = x.shape[0]
        if passes is not None and self.ls_umma_supported(k, passes):
            mt = (b.H * b.W + 127) // 128
            S = self._splits(mt, B, n_sel, waves=2, min_images=128)
            P = self._partials(tag, S, B)
            _lib.check(self.lib.cds_ls_partials_umma(
                _lib.ptr(x), B, b.C, b.H, b.W, k, _lib.ptr(beta), _lib.ptr(b.flat16()), b.strip8()[2], _lib.ptr(b.ls_norms(k)),
                _lib.ptr(idx), _lib.ptr(logw), n_sel, S, passes, _lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc),
                _lib.stream_ptr()), "cds_ls_partials_umma")
            self.launches += 1
            return P
        rows = bool(self.lib.cds_ls_rows_supported(b.C, b.H, b.W, k))
        sb = 2 if rows else 4                                   # samples per CTA of the two kernels
        S = int(min(n_sel, max(1, (2 * sm_count(self.device)) // ((B + sb - 1) // sb))))
        P = self._partials(tag, S, B)
        fn = self.lib.cds_ls_rows_partials if rows else self.lib.cds_ls_partials
        _lib.check(fn(_lib.ptr(x), B, b.C, b.H, b.W, k, _lib.ptr(beta), _lib.ptr(b.images),
                                            _lib.ptr(idx), _lib.ptr(logw), n_sel, S, _lib.ptr(P.m), _lib.ptr(P.l),
                                            _lib.ptr(P.acc), _lib.stream_ptr()), "cds_ls_partials")
        self.launches += 1
        return P

    def edge_supported(self, k):
        b = self.bank
        return self.use_tensor_cores and bool(self.lib.cds_bbels_edge_supported(b.C, b.H, b.W, k))

    def edge_umma_supported(self, k, passes):
        """Tensor-core edge bands: single exact bank plane, square images, a geometry that fits shared memory."""
        b = self.bank
        return (self.use_tensor_cores and self.edge_variant != "simt" and b.H == b.W
                and b.strip8()[1] is None and self.lib.cds_bbels_edge_umma_smem_bytes(b.C, b.H, b.W, k, passes) > 0)

    def edge_partials(self, x, beta, k, sel, tag="edge", passes=2):
        """bbELS edge bands: csrc/bbels_edge_umma.cu (tcgen05, `passes` as for the centre) where the geometry fits, else the
        exact fp32 SIMT kernel csrc/bbels_edge.cu; writes the edge pixels of the partials only."""
        idx, logw, n_sel = sel
        if n_sel == 0:
            return self._empty_shard(tag, x.shape[0])
        b = self.bank
        B = x.shape[0]
        if self.edge_umma_supported(k, passes):
            d = k // 2
            mt = (((b.H - 2 * d + 7) // 8) * d + 15) // 16           # M tiles per band (bbels_edge_umma.cu: EdgeGeom::MT)
            S = self._splits(4 * mt, B, n_sel, waves=1, min_images=32)
            P = self._partials(tag, S, B)
            _lib.check(self.lib.cds_bbels_edge_partials_umma(
                _lib.ptr(x), B, b.C, b.H, b.W, k, _lib.ptr(beta), _lib.ptr(b.edge_plane()), b.strip8()[2],
                _lib.ptr(b.edge_norms(k)), _lib.ptr(idx), _lib.ptr(logw), n_sel, S, passes, _lib.ptr(P.m), _lib.ptr(P.l),
                _lib.ptr(P.acc), _lib.stream_ptr()), "cds_bbels_edge_partials_umma")
            self.launches += 1
            return P
        S = int(max(1, min(n_sel // 8, (4 * sm_count(self.device)) // (4 * B))))
        P = self._partials(tag, S, B)
        _lib.check(self.lib.cds_bbels_edge_partials(_lib.ptr(x), B, b.C, b.H, b.W, k, _lib.ptr(beta), _lib.ptr(b.images),
                                                    _lib.ptr(idx), _lib.ptr(logw), n_sel, S, _lib.ptr(P.m), _lib.ptr(P.l),
                                                    _lib.ptr(P.acc), _lib.stream_ptr()), "cds_bbels_edge_partials")
        self.launches += 1
        return P

    def umma_partials(self, pad, x, beta, k, sel, passes, dbg=None, tag="umma", a_over_beta=None, window=None):
        """window = (i0, j0, rows, cols): only the query tiles covering that pixel window are launched (bbELS centre)."""
        idx, logw, n_sel = sel
        if n_sel == 0:
            return self._empty_shard(tag, x.shape[0])
        b = self.bank
        B = x.shape[0]
        hi, lo, scale = b.strip8()
        pn = b.norm_plane(k)
        qi0, qj0, qrows, qcols = window if window is not None else (0, 0, b.H, b.W)
        tiles = ((qrows + 15) // 16) * ((qcols + 7) // 8)
        S = self._splits(tiles, B, n_sel)
        P = self._partials(tag, S, B)
        planes = 1 if lo is None else 2
        variant = _lib.ELS_VARIANT[self.els_variant]
        if variant == 2 and (dbg is not None or not self.lib.cds_els_umma_pv_supported(b.C, b.H, b.W, k, passes, planes)):
            variant = 1                                            # "pv" means: wherever the geometry allows it
        if variant == 0 and n_sel < 400 * S:
            # the P.V epilogue has a start-up cost per CTA (the first tile runs on the exact path, ~11 re-basings of the
            # reference): with fewer than ~400 images per CTA -- small banks, or the headline bank sharded over 4-8 GPUs --
            # the FMA epilogue is faster (6 250-image bank, 70 images per CTA: 18.8 vs 20.7 ms per step)
            variant = 1
        if variant == 0 and a_over_beta is not None and a_over_beta > 100.0:
            # at the lowest noise levels most 16-column chunks carry no weight at all (77 % at a/beta = 165 on the headline
            # trajectory): the FMA-pipe epilogue skips them outright, the P.V epilogue still stores and contracts zeros
            variant = 1
        rows = b.rows8() if (k > 8 and k % 8) else None            # mixed K layout for the trailing k % 8 patch rows
        _lib.check(self.lib.cds_els_partials_umma_window(
            _lib.PAD[pad], _lib.ptr(x), B, b.C, b.H, b.W, k, _lib.ptr(beta), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(rows),
            scale, _lib.ptr(pn), _lib.ptr(idx), _lib.ptr(logw), n_sel, S, passes, variant, qi0, qj0, qrows, qcols,
            _lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), _lib.ptr(dbg), _lib.stream_ptr()), "cds_els_partials_umma_window")
        self.launches += 1
        return P

    def combine(self, P):
        """Merge the S slices into slice 0 (in place), then across ranks when the bank is sharded."""
        if self.group is not None:
            from .distributed import allgather_combine
            return allgather_combine(self, P, P.S)
        if P.S > 1:
            _lib.check(self.lib.cds_combine(_lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), P.S, P.B, P.C, P.HW,
                                            _lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), _lib.stream_ptr()),
                       "cds_combine")
            self.launches += 1
        return P

    def finish(self, P, x, beta, mu, score, region=0, d=0, step=None):
        """Fused tail of one evaluation (cds_finish): merge of the CTA slices (and, with a process group, of the all-gathered
        per-rank records), mu / score, and -- when `step` is given -- the sampler update of x in place.  `step` is a dict
        with device tensors cx, cmu [B] and optionally sigma [B] plus noise [B,C,H,W] or seed_offset (uint64 [2]) and the
        int `index` of the step (Philox counter)."""
        b = self.bank
        st = step or {}
        sigma, noise, so = st.get("sigma"), st.get("noise"), st.get("seed_offset")
        args = (x.shape[0], b.C, b.H, b.W, region, d, _lib.ptr(x), _lib.ptr(beta), _lib.ptr(mu), _lib.ptr(score),
                _lib.ptr(st.get("cx")), _lib.ptr(st.get("cmu")), _lib.ptr(sigma), _lib.ptr(noise), _lib.ptr(so),
                int(st.get("index", 0)), _lib.stream_ptr())
        if self.group is not None:
            from .distributed import allgather_packed
            gathered, world = allgather_packed(self, P, P.S)
            _lib.check(self.lib.cds_finish(_lib.ptr(gathered), None, None, 1, world, *args), "cds_finish")
        else:
            _lib.check(self.lib.cds_finish(_lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), 0, P.S, *args), "cds_finish")
        self.launches += 1

    def finalize(self, P, x, beta, mu, score, region=0, d=0):
        b = self.bank
        _lib.check(self.lib.cds_finalize(_lib.ptr(x), _lib.ptr(beta), _lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc),
                                         x.shape[0], b.C, b.H, b.W, region, d, _lib.ptr(mu), _lib.ptr(score),
                                         _lib.stream_ptr()), "cds_finalize")
        self.launches += 1

    def ddim_step(self, x, mu, c_x, c_mu):
        B = x.shape[0]
        _lib.check(self.lib.cds_ddim_step(_lib.ptr(x), _lib.ptr(mu), _lib.ptr(c_x), _lib.ptr(c_mu), B,
                                          x.numel() // B, _lib.stream_ptr()), "cds_ddim_step")
        self.launches += 1

    # ---- one score evaluation ------------------------------------------------------------------
    def evaluate(self, kind, x, beta, k, sel, query_pad=None, mu=None, score=None, beta_min=None, sel_ls=None, step=None):
        """x [B,C,H,W] fp32 on device, beta [B] fp32 on device.  Writes mu and/or score when given; with `step` (see
        finish) x is advanced in place by the sampler update.  Returns (mu, score)."""
        b = self.bank
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert tuple(x.shape[1:]) == (b.C, b.H, b.W), f"x {tuple(x.shape)} does not match bank {(b.C, b.H, b.W)}"
        if x.device != self.device:
            raise RuntimeError(f"x lives on {x.device} but the bank of this engine on {self.device}")
        # the raw launches below go to the CURRENT device's stream: make the engine's device current for their duration
        with torch.cuda.device(self.device):
            return self._evaluate(kind, x, beta, k, sel, query_pad, mu, score, beta_min, sel_ls, step)

    def _evaluate(self, kind, x, beta, k, sel, query_pad, mu, score, beta_min, sel_ls, step):
        b = self.bank
        if kind == "IS":                          # whole-image window: the LS kernel with k = 2*max(H,W)-1; the
            kind, k = "LS", 2 * max(b.H, b.W) - 1  # reference IS ignores k altogether (idealscore.py:583, **kwargs)
        if k is None or k % 2 == 0 or k < 1:
            raise ValueError(f"kernel size must be odd and positive, got {k}")
        if beta_min is None:
            beta_min = float(beta.min())         # host sync; the machine passes beta_min explicitly
        if kind == "bbELS" and k >= b.H:          # idealscore.py:163-164: delegate to the internal LS module
            kind, sel = "LS", (sel_ls if sel_ls is not None else sel)
        if kind == "LS":
            # use_tensor_cores=False asks for the generic exact-fp32 kernel; the streaming LS kernels are fp32 SIMT as
            # well and take over where the generic one cannot hold the padded planes (IS on images above 32 pixels)
            passes = self.passes_for(k, beta_min)
            if self.ls_umma_supported(k, passes):
                P = self.ls_partials(x, beta, k, sel, passes=passes)
            elif self.ls_supported(k) and (self.use_tensor_cores or not self.simt_supported(k)):
                P = self.ls_partials(x, beta, k, sel)
            else:
                P = self.simt_partials("LS", "zeros", x, beta, k, sel)
            self.finish(P, x, beta, mu, score, step=step)
        elif kind == "ELS":
            if k > b.H or k > b.W:
                raise ValueError(f"ELS needs kernel size <= image size, got k={k} for {b.H}x{b.W}")
            pad = query_pad or "circular"
            passes = self.passes_for(k, beta_min)
            if self.umma_supported(k, passes):
                P = self.umma_partials(pad, x, beta, k, sel, passes, a_over_beta=self._a_over_beta(beta_min))
            else:
                P = self.simt_partials("ELS", pad, x, beta, k, sel)
            self.finish(P, x, beta, mu, score, step=step)
        elif kind == "bbELS":
            passes = self.passes_for(k, beta_min)
            d = k // 2
            # every partials kernel reads x, so all of them run before the first finish (which may advance x in place);
            # the regions are disjoint sets of pixels, each finish touches only its own
            if self.umma_supported(k, passes):
                has_edges = (self.edge_umma_supported(k, passes) or self.edge_supported(k)) and self.ls_supported(k)
                # only the centre pixels of this kernel's output are used: launch just the query tiles that cover them
                window = (d, d, b.H - 2 * d, b.W - 2 * d) if self.centre_window else None
                Pc = self.umma_partials("zeros", x, beta, k, sel, passes, a_over_beta=self._a_over_beta(beta_min),
                                        window=window)
                if has_edges:
                    # edge bands: dedicated kernel; corners see only their own location = the LS kernel
                    Pe = self.edge_partials(x, beta, k, sel, passes=passes)
                    Pk = self.ls_partials(x, beta, k, sel, tag="corner")
                    self.finish(Pc, x, beta, mu, score, region=1, d=d, step=step)
                    self.finish(Pe, x, beta, mu, score, region=4, d=d, step=step)
                    self.finish(Pk, x, beta, mu, score, region=3, d=d, step=step)
                else:
                    Pb = self.simt_partials("bbELS", "zeros", x, beta, k, sel, region=2, tag="border")
                    self.finish(Pc, x, beta, mu, score, region=1, d=d, step=step)
                    self.finish(Pb, x, beta, mu, score, region=2, d=d, step=step)
            else:
                P = self.simt_partials("bbELS", "zeros", x, beta, k, sel)
                self.finish(P, x, beta, mu, score, step=step)
        else:
            raise ValueError(f"unknown score module kind {kind!r}")
        return mu, score
