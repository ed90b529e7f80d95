"""world_size-2 gloo tests (CPU) of the bank-sharding plumbing: interleaved shards + all-gather of the
(max, sum-exp, weighted-sum) partials + log-sum-exp merge reproduce the un-sharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import score_oracle as so
from convolutional_diffusion_b200 import selection as sel
from convolutional_diffusion_b200.distributed import gather_partials


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials_ls(x, bank, beta, k, logw):
    """(m, l, acc) in log2 units for LS on a bank slice -- float64 restatement used only by this test."""
    from numpy.lib.stride_tricks import sliding_window_view
    a = np.sqrt(1 - beta)
    d = k // 2
    e = ((x[None] - a * bank) ** 2).sum(1)
    e = np.pad(e, ((0, 0), (d, d), (d, d)))
    box = sliding_window_view(e, (k, k), axis=(1, 2)).sum(axis=(-1, -2))
    t = (-box / (2 * beta) + logw[:, None, None]) * np.log2(np.e)
    m = t.max(0)
    p = np.exp2(t - m[None])
    return m, p.sum(0), (p[:, None] * bank).sum(0)


def _combine(m, l, acc):
    M = m.max(0).values
    w = torch.exp2(m - M[None])
    return M, (l * w).sum(0), (acc * w[:, None, :]).sum(0)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n, c, h, k, beta, label, bs = 37, 2, 10, 3, 0.35, 1, 8
    bank = rng.uniform(-1, 1, (n, c, h, h))
    labels = rng.integers(0, 3, n)
    x = rng.normal(size=(c, h, h))
    idx, logw = sel.select("LS", labels, label, bs, None)
    owner = sel.assign_ranks(labels, world)              # bank sharding by image ownership: each rank holds ~N/world images
    my_idx, my_logw = sel.shard(idx, logw, rank, world, owner)
    assert set(my_idx.tolist()) <= set(np.nonzero(owner == rank)[0].tolist())
    m, l, acc = _partials_ls(x, bank[my_idx], beta, k, my_logw)
    B, HW = 1, h * h
    gm, gl, gacc = gather_partials(torch.from_numpy(m).reshape(B, HW), torch.from_numpy(l).reshape(B, HW),
                                   torch.from_numpy(acc).reshape(B, c, HW))
    assert gm.shape == (world, B, HW) and gacc.shape == (world, B, c, HW)
    M, L, A = _combine(gm[:, 0], gl[:, 0], gacc[:, 0])
    mu = (A / L[None]).reshape(c, h, h).numpy()
    ref = so.ls_mu(x, bank[idx], beta, k, logw)
    out[rank] = float(np.max(np.abs(mu - ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_partials_merge_to_full_result():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r] < 1e-10, out[r]


def _worker_order(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from convolutional_diffusion_b200.distributed import shared_order
    torch.manual_seed(100 + rank)                       # ranks with DIFFERENT global RNG states
    mine = sel.dataloader_shuffle_order(53)
    order = shared_order(mine, dist.group.WORLD)
    out[rank] = (mine.tolist(), order.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_shuffled_selection_uses_one_order_on_all_ranks():
    """Bank sharding + a shuffled DataLoader order (LS, shuffle=True): ranks seeded differently draw different permutations;
    the shared order is rank 0's on every rank, so the shards partition one selection (no image twice, none dropped)."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_order, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] != out[1][0]                        # the local draws differ ...
    assert out[0][1] == out[1][1] == out[0][0]           # ... the shared order is rank 0's
    labels = np.random.default_rng(1).integers(0, 3, 53)
    owner = sel.assign_ranks(labels, world)
    idx, logw = sel.select("LS", labels, 1, 8, None, np.asarray(out[0][1]))
    parts = [sel.shard(idx, logw, r, world, owner)[0] for r in range(world)]
    assert sorted(np.concatenate(parts).tolist()) == sorted(idx.tolist())
