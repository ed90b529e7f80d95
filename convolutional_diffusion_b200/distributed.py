"""Bank sharding across the GPUs of one box.

The softmax over the bank is associative under (max, sum-exp, weighted-sum), so the training images are split
across ranks (interleaved, selection.shard) with NO collective on the data path; every rank reduces its slice to
partials for all B*H*W queries and the only exchange is one small all-gather of those partials per score
evaluation ((2+C)*B*H*W floats per rank: 20 KB at B=1 CIFAR), followed by the same log-sum-exp merge kernel
that merges CTA splits (cds_combine).  mu and the DDIM update are then computed redundantly on every rank, which
keeps x bit-identical across ranks (the merge order is rank order everywhere).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def gather_partials(m, l, acc, group=None):
    """all-gather of one rank's merged partials: m,l [B,HW], acc [B,C,HW] -> [W,B,HW], [W,B,HW], [W,B,C,HW].
    Works on any backend (nccl on the GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    out = []
    for t in (m, l, acc):
        t = t.contiguous()
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t, group=group)          # rank-major concatenation along dim 0
        out.append(g.view((world,) + tuple(t.shape)))
    return out


def allgather_combine(engine, P, local_S):
    """Merge the local_S CTA slices of this rank into a packed [m | l | acc] record, all-gather the records (ONE
    collective, (2+C)*B*HW floats per rank) and merge them in rank order into slice 0 of P, identically on all ranks.
    Buffers are cached so the sequence is CUDA-graph capturable."""
    group = engine.group
    world = dist.get_world_size(group)
    key = ("packed", P.B)
    if key not in engine._buf:
        n = P.B * (2 + P.C) * P.HW
        engine._buf[key] = (torch.empty(n, dtype=torch.float32, device=P.m.device),
                            torch.empty(world * n, dtype=torch.float32, device=P.m.device))
    mine, gathered = engine._buf[key]
    bhw = P.B * P.HW
    _lib.check(engine.lib.cds_combine(_lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), local_S, P.B, P.C, P.HW,
                                      _lib.ptr(mine), _lib.ptr(mine[bhw:]), _lib.ptr(mine[2 * bhw:]),
                                      _lib.stream_ptr()), "cds_combine")
    dist.all_gather_into_tensor(gathered, mine, group=group)
    _lib.check(engine.lib.cds_combine_packed(_lib.ptr(gathered), world, P.B, P.C, P.HW, _lib.ptr(P.m), _lib.ptr(P.l),
                                             _lib.ptr(P.acc), _lib.stream_ptr()), "cds_combine_packed")
    engine.launches += 2
    return P


def allgather_packed(engine, P, local_S):
    """First half of allgather_combine for the fused tail (cds_finish): merges this rank's CTA slices into its packed
    record and all-gathers the records; returns (gathered [world * B*(2+C)*HW], world).  The merge across ranks then happens
    inside cds_finish, in rank order on every rank (x stays bit-identical across ranks)."""
    group = engine.group
    world = dist.get_world_size(group)
    key = ("packed", P.B)
    if key not in engine._buf:
        n = P.B * (2 + P.C) * P.HW
        engine._buf[key] = (torch.empty(n, dtype=torch.float32, device=P.m.device),
                            torch.empty(world * n, dtype=torch.float32, device=P.m.device))
    mine, gathered = engine._buf[key]
    bhw = P.B * P.HW
    _lib.check(engine.lib.cds_combine(_lib.ptr(P.m), _lib.ptr(P.l), _lib.ptr(P.acc), local_S, P.B, P.C, P.HW,
                                      _lib.ptr(mine), _lib.ptr(mine[bhw:]), _lib.ptr(mine[2 * bhw:]),
                                      _lib.stream_ptr()), "cds_combine")
    dist.all_gather_into_tensor(gathered, mine, group=group)
    engine.launches += 1
    return gathered, world


def shared_order(order, group, device=None):
    """The DataLoader visiting order every rank must use when the bank is sharded and the order is drawn from the global
    torch RNG (LS hard-codes shuffle=True, idealscore.py:489): each rank has consumed its own two RNG draws, as one
    reference evaluation does; rank 0's permutation is broadcast so that all ranks split the same selection."""
    import numpy as np
    t = torch.from_numpy(np.array(order, dtype=np.int64))          # a copy: the broadcast must not overwrite the caller's draw
    if dist.get_backend(group) == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0), group=group)
    return t.cpu().numpy()


def init_from_env():
    """One process per GPU (torchrun): returns (rank, world, local_rank) and initialises NCCL if world > 1."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, world, local


def shutdown(timeout_s=30.0):
    """Orderly teardown of a multi-rank run whose trajectories were captured in CUDA graphs: the graphs reference the NCCL
    communicator (the all-gathers are captured nodes), and destroying the communicator while such graphs are alive -- or
    leaving both to interpreter shutdown in arbitrary order -- can block forever.  Callers first drop their graphs
    (ScheduledScoreMachine.release_graphs), then call this: drain the device, agree that every rank is done, destroy
    the group.  A watchdog turns a blocked destructor into a hard exit instead of a hung job (it has never fired in the
    runs of round 2; it is there because a hang on a shared GPU box is worse than a missing destructor)."""
    import os
    import threading
    if not dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    done = threading.Event()

    def _watchdog():
        if not done.wait(timeout_s):
            import sys
            sys.stderr.write("convolutional_diffusion_b200.distributed.shutdown: process group teardown blocked, exiting hard\n")
            sys.stderr.flush()
            os._exit(0)

    threading.Thread(target=_watchdog, daemon=True).start()
    dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.destroy_process_group()
    done.set()
