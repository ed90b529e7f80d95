"""The training-patch bank, resident in HBM.

The reference re-reads and re-transforms the whole dataset through a DataLoader on every score evaluation
(`/root/reference/src/utils/idealscore.py:184,430,521`) and copies each batch host->device inside the hot loop
(:441).  Here the images are uploaded once and kept in two layouts:

  images  fp32 planar [N,C,H,W]            exact values; SIMT kernels, patch norms, LS stream
  strip8  fp16 [N,C,H,W,8] (+ residual)    tensor-core stream: one 16-byte granule = 8 vertically adjacent
                                           pixels, so overlapping k x k patches alias the same bytes
                                           (implicit im2col, see csrc/els_umma.cu)

8-bit image datasets normalised with mean 0.5 / std 0.5 (`src/utils/data.py:63-70`) are odd integers / 255,
which fp16 represents exactly after scaling by 255 -> a single fp16 plane carries the bank losslessly.
Other banks are scaled by 256 and get a second plane holding the fp16 rounding residual.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .selection import assign_ranks, select, shard


def dataset_to_tensors(dataset):
    """Materialises a map-style dataset of (image [C,H,W], label) pairs (the protocol of
    idealscore.py:142,390,489) into (images [N,C,H,W] float32, labels [N] int64) on the host."""
    if isinstance(dataset, (tuple, list)) and len(dataset) == 2 and torch.is_tensor(dataset[0]):
        return dataset[0].float().contiguous(), torch.as_tensor(dataset[1]).long()
    imgs = getattr(dataset, "images", None)
    labs = getattr(dataset, "labels", None)
    if torch.is_tensor(imgs) and labs is not None and imgs.dim() == 4:
        return imgs.float().contiguous(), torch.as_tensor(labs).long()
    if isinstance(dataset, torch.utils.data.Subset):
        base_i, base_l = dataset_to_tensors(dataset.dataset)
        sel = torch.as_tensor(list(dataset.indices), dtype=torch.long)
        return base_i[sel].contiguous(), base_l[sel]
    if isinstance(dataset, torch.utils.data.TensorDataset) and len(dataset.tensors) == 2:
        return dataset.tensors[0].float().contiguous(), dataset.tensors[1].long()
    n = len(dataset)
    first, _ = dataset[0]
    images = torch.empty((n,) + tuple(first.shape), dtype=torch.float32)
    labels = torch.empty(n, dtype=torch.long)
    for i in range(n):
        im, lab = dataset[i]
        images[i] = im
        labels[i] = int(lab)
    return images, labels


class PatchBank:
    """rank / world: bank sharding across the GPUs of one box.  The rank keeps (uploads, packs, streams) only the images it
    owns under selection.assign_ranks -- about N/world of them -- while the labels of the whole dataset stay on the host,
    because the reference's DataLoader bookkeeping (which images take part, with which weight) is defined on the whole
    dataset."""

    def __init__(self, images, labels, device=None, rank=0, world=1):
        self.lib = _lib.load()                                   # raises if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("PatchBank needs a CUDA device: the score machines have no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError(f"PatchBank device must be CUDA, got {self.device}")
        if self.device.index is None:                            # always a concrete ordinal: tensors report cuda:N
            self.device = torch.device("cuda", torch.cuda.current_device())
        assert images.dim() == 4, "bank images must be [N,C,H,W]"
        self.labels = np.asarray(torch.as_tensor(labels).cpu()).astype(np.int64).reshape(-1)
        self.N = int(images.shape[0])                            # images of the whole dataset (selection bookkeeping)
        assert self.labels.shape[0] == self.N
        self.rank, self.world = int(rank), int(world)
        self.owner = assign_ranks(self.labels, self.world)
        if self.world > 1:
            mine = np.nonzero(self.owner == self.rank)[0]
            self.local_of = np.full(self.N, -1, dtype=np.int64)  # dataset index -> index in this rank's device arrays
            self.local_of[mine] = np.arange(mine.shape[0])
            images = images[torch.from_numpy(mine)]
        else:
            self.local_of = None
        self.images = images.to(self.device, torch.float32).contiguous()
        self.N_local, self.C, self.H, self.W = self.images.shape
        self._strip = None
        self._rows8 = None
        self._pnorm = {}
        self._nplane = {}
        self._eplane = None
        self._enorms = {}
        self._flat16 = None
        self._lsnorms = {}
        self._sel = {}

    # ---- layouts -------------------------------------------------------------------------------
    def strip8(self):
        """(hi, lo or None, scale): the fp16 strip layout for the tensor-core kernel."""
        if self._strip is None:
            with torch.cuda.device(self.device):
                v = self.images * 127.5 + 127.5                    # back to the 0..255 grid of ToTensor()
                exact8 = bool(((v - v.round()).abs().max() < 1e-3).item()) and bool((v.min() > -0.5).item()) \
                    and bool((v.max() < 255.5).item())
                scale = 255.0 if exact8 else 256.0
                n = self.N_local * self.C * self.H * self.W
                hi = torch.empty(n * 8, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_strip8(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W,
                                                    scale, 0, _lib.ptr(hi), _lib.stream_ptr()), "cds_pack_strip8")
                lo = None
                if not exact8:
                    lo = torch.empty(n * 8, dtype=torch.float16, device=self.device)
                    _lib.check(self.lib.cds_pack_strip8(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W,
                                                        scale, 1, _lib.ptr(lo), _lib.stream_ptr()), "cds_pack_strip8")
                self._strip = (hi, lo, scale)
        return self._strip

    def rows8(self):
        """fp16 plane of 8-pixel HORIZONTAL strips (same scale as strip8), or None for a two-plane bank: lets the
        tensor-core kernel contract the trailing k % 8 patch rows without padding them to a block of 8."""
        hi, lo, scale = self.strip8()
        if lo is not None:
            return None
        if self._rows8 is None:
            with torch.cuda.device(self.device):
                out = torch.empty(self.N_local * self.C * self.H * self.W * 8, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_strip8(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W,
                                                    scale, 2, _lib.ptr(out), _lib.stream_ptr()), "cds_pack_strip8")
                self._rows8 = out
        return self._rows8

    def patch_norms(self, k):
        """||p||^2 of every valid k x k x C patch, [N, (H-k+1)*(W-k+1)] fp32 (computed once per k)."""
        if k not in self._pnorm:
            with torch.cuda.device(self.device):
                out = torch.empty(self.N_local * (self.H - k + 1) * (self.W - k + 1), dtype=torch.float32, device=self.device)
                _lib.check(self.lib.cds_patch_norms(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W, k,
                                                    _lib.ptr(out), _lib.stream_ptr()), "cds_patch_norms")
                self._pnorm[k] = out
        return self._pnorm[k]

    def norm_plane(self, k):
        """fp16 [N,H,W,8] norm plane of kernel size k for the tensor-core kernel (computed once per k)."""
        if k not in self._nplane:
            with torch.cuda.device(self.device):
                out = torch.empty(self.N_local * self.H * self.W * 8, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_norm_plane(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W, k,
                                                        _lib.ptr(out), _lib.stream_ptr()), "cds_pack_norm_plane")
                self._nplane[k] = out
        return self._nplane[k]

    def edge_plane(self):
        """fp16 plane of the four border bands in one orientation (8-pixel granules ACROSS the band, k independent) for the
        tensor-core bbELS edge kernel; same scale as strip8.  None for a two-plane bank."""
        hi, lo, scale = self.strip8()
        if lo is not None:
            return None
        if self._eplane is None:
            with torch.cuda.device(self.device):
                n = int(self.lib.cds_edge_plane_halves(self.N_local, self.C, self.H))
                out = torch.empty(n, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_edge_plane(_lib.ptr(self.images), self.N_local, self.C, self.H, scale,
                                                        _lib.ptr(out), _lib.stream_ptr()), "cds_pack_edge_plane")
                self._eplane = out
        return self._eplane

    def edge_norms(self, k):
        """fp16 squared norms of the truncated border patches of kernel size k as K granules (computed once per k)."""
        if k not in self._enorms:
            with torch.cuda.device(self.device):
                n = int(self.lib.cds_edge_norms_halves(self.N_local, self.H, k))
                out = torch.empty(n, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_edge_norms(_lib.ptr(self.images), self.N_local, self.C, self.H, k,
                                                        _lib.ptr(out), _lib.stream_ptr()), "cds_pack_edge_norms")
                self._enorms[k] = out
        return self._enorms[k]

    def flat16(self):
        """fp16 [N][HWp] flattened single-channel images (pixel * scale, rows padded to whole 16-byte granules) for the
        tensor-core LS kernel; None for a two-plane or multi-channel bank."""
        hi, lo, scale = self.strip8()
        if lo is not None or self.C != 1:
            return None
        if self._flat16 is None:
            with torch.cuda.device(self.device):
                n = int(self.lib.cds_ls_plane_elems(self.N_local, self.H, self.W))
                out = torch.empty(n, dtype=torch.float16, device=self.device)
                _lib.check(self.lib.cds_pack_flat16(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W, scale,
                                                    _lib.ptr(out), _lib.stream_ptr()), "cds_pack_flat16")
                self._flat16 = out
        return self._flat16

    def ls_norms(self, k):
        """fp32 [N][128*ceil(HW/128)] zero-padded k x k window sums of T^2 (computed once per k) for the tensor-core LS kernel."""
        if k not in self._lsnorms:
            with torch.cuda.device(self.device):
                n = int(self.lib.cds_ls_norms_elems(self.N_local, self.H, self.W))
                out = torch.empty(n, dtype=torch.float32, device=self.device)
                _lib.check(self.lib.cds_pack_ls_norms(_lib.ptr(self.images), self.N_local, self.C, self.H, self.W, k,
                                                      _lib.ptr(out), _lib.stream_ptr()), "cds_pack_ls_norms")
                self._lsnorms[k] = out
        return self._lsnorms[k]

    def ls_bytes_per_pixel(self):
        """Bytes per bank pixel the LS kernels stream from HBM (algorithmic bytes of the roofline)."""
        return 4

    # ---- selection -----------------------------------------------------------------------------
    def selection(self, kind, label, batch_size, max_samples, order=None):
        """(idx int32 device, logw fp32 device, n_sel) for one score evaluation: the reference's DataLoader bookkeeping on the
        whole dataset, then -- with a sharded bank -- the part owned by this rank, as indices into its device arrays.
        Cached per key when the visiting order is the identity."""
        key = (kind, label, batch_size, max_samples) if order is None else None
        if key is not None and key in self._sel:
            return self._sel[key]
        idx, logw = select(kind, self.labels, label, batch_size, max_samples, order)
        if idx.shape[0] == 0:
            raise RuntimeError(f"no bank image selected (kind={kind}, label={label}, max_samples={max_samples})")
        if self.world > 1:
            idx, logw = shard(idx, logw, self.rank, self.world, self.owner)
            idx = self.local_of[idx]
        out = (torch.from_numpy(idx.astype(np.int32)).to(self.device),
               torch.from_numpy(logw.astype(np.float32)).to(self.device), int(idx.shape[0]))
        if key is not None:
            self._sel[key] = out
        return out
