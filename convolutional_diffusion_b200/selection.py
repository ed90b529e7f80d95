"""Host-side replay of the reference's DataLoader bookkeeping: which bank images take part in one score
evaluation and with which weight.

The reference streams the training set in DataLoader batches and combines batches with an online softmax.
Three quirks decide the result and are reproduced here as an index list + per-image log-weight, so that the
CUDA kernels see one flat, already filtered bank (`/root/reference/src/utils/idealscore.py`):

  ELS   :430-444  running count of *pre-filter* images; `break` once it exceeds max_samples.
        :470-471  torch.mean over the patches of each batch -> every image of a batch that kept n_b images
                  after the label filter carries weight 1/n_b  (log-weight -log n_b).
  LS    :521-535  running count of *post-filter* images, same `break`; :553-554 the same per-batch mean.
        :489      shuffle=True is hard-coded.
  bbELS :184-193  counter advances by batch_size per visited batch, checked before the batch; :336-368 plain
                  torch.sum -> weight 1.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["select", "shard", "assign_ranks", "dataloader_shuffle_order"]


def select(kind, labels, label, batch_size, max_samples, order=None):
    """Returns (idx int64 [n_sel], logw float64 [n_sel]) in visiting order."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    order = np.arange(n, dtype=np.int64) if order is None else np.asarray(order, dtype=np.int64)
    if kind == "IS":                      # idealscore.py:600-614: the same bookkeeping as LS
        kind = "LS"
    if kind not in ("LS", "ELS", "bbELS"):
        raise ValueError(f"unknown score module kind {kind!r}")
    batch_size = int(batch_size)
    keep = np.ones(n, dtype=bool) if label is None else (labels[order] == int(label))
    nb = (n + batch_size - 1) // batch_size
    starts = np.arange(nb) * batch_size
    sizes = np.minimum(batch_size, n - starts)
    kept = np.add.reduceat(keep.astype(np.int64), starts)          # images surviving the filter, per batch
    # number of leading batches that are processed before the max_samples `break`
    if max_samples is None:
        nvisit = nb
    elif kind == "ELS":
        seen = np.cumsum(sizes)                                     # counted before the filter
        nvisit = int(np.searchsorted(seen > max_samples, True))
    elif kind == "LS":
        seen = np.cumsum(kept)                                      # counted after the filter; empty batches skipped
        nvisit = int(np.searchsorted(seen > max_samples, True))
    else:
        seen_before = np.arange(nb) * batch_size                    # q before the batch; q += batch_size after
        nvisit = int(np.searchsorted(seen_before > max_samples, True))
    pos = np.nonzero(keep[:min(n, nvisit * batch_size)])[0]
    idx = order[pos]
    if kind == "bbELS":
        logw = np.zeros(len(idx))
    else:
        with np.errstate(divide="ignore"):
            logw = -np.log(kept[pos // batch_size].astype(np.float64))
    return idx.astype(np.int64), logw


def assign_ranks(labels, world):
    """Owner rank of every bank image for bank sharding: the q-th image of its class (in dataset order) belongs to rank
    q % world.  The assignment depends on the image only -- not on the selection -- so a rank uploads and packs just the
    ~N/world images it owns, and every selection (any label, max_samples, visiting order) is split into the subsets
    of its images owned by each rank: balanced per class, hence balanced for conditional and unconditional runs alike."""
    labels = np.asarray(labels)
    owner = np.zeros(labels.shape[0], dtype=np.int64)
    if world <= 1:
        return owner
    order = np.argsort(labels, kind="stable")
    sorted_labels = labels[order]
    first = np.r_[0, np.nonzero(np.diff(sorted_labels))[0] + 1]            # start of every class run
    start_of = np.repeat(first, np.diff(np.r_[first, labels.shape[0]]))
    owner[order] = (np.arange(labels.shape[0]) - start_of) % world
    return owner


def shard(idx, logw, rank, world, owner=None):
    """The part of a selection that lives on `rank`: its images owned by that rank (assign_ranks), in selection order; the
    log-weights travel with the images.  No data-path collective is needed -- every rank reduces its part to
    (max, sum-exp, weighted-sum) partials.  Without `owner` the selection is dealt out round robin (idx[rank::world])."""
    if owner is None:
        return idx[rank::world], logw[rank::world]
    keep = owner[idx] == rank
    return idx[keep], logw[keep]


def dataloader_shuffle_order(n):
    """Visiting order of `DataLoader(dataset, shuffle=True)` for the current torch global RNG state
    (LS hard-codes shuffle=True, idealscore.py:489).  torch's iterator first draws the loader base seed,
    then RandomSampler draws its own seed from the default generator and permutes with a fresh generator."""
    import torch
    torch.empty((), dtype=torch.int64).random_()                    # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())  # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).numpy()
