"""Multi-GPU check (run under torchrun, one rank per GPU): the bank-sharded ELS / bbELS / LS evaluation and a full
trajectory agree with the un-sharded ones computed on the same rank, and x stays bit-identical across ranks.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dist_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import (LocalEquivScoreModule, LocalEquivBordersScoreModule, LocalScoreModule,  # noqa: E402
                                           ScheduledScoreMachine, cosine_noise_schedule)
from convolutional_diffusion_b200.distributed import init_from_env  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402


def main():
    rank, world, local = init_from_env()
    dev = torch.device("cuda", local)
    bank, labels = synthetic_bank(3000, 3, 32, nlabels=10, seed=0)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(3)).to(dev)
    worst = 0.0
    for cls, k, t in ((LocalEquivScoreModule, 7, 0.4), (LocalEquivBordersScoreModule, 9, 0.6), (LocalScoreModule, 5, 0.3)):
        kw = dict(kernel_size=k, batch_size=64, schedule=cosine_noise_schedule)
        full = cls((bank, labels), **kw)
        shard = cls((bank, labels), process_group=dist.group.WORLD, **kw)
        for label in (None, torch.tensor([4])):
            torch.manual_seed(11)        # LS reshuffles its DataLoader per call (idealscore.py:489): same draws for both
            s0 = full(torch.full((2,), t), x, label=label, device=dev)
            torch.manual_seed(11)
            s1 = shard(torch.full((2,), t), x, label=label, device=dev)
            err = float((s0 - s1).abs().max())
            worst = max(worst, err)
            assert err < 1e-3, (cls.__name__, label, err)
    # single-channel bank: LS goes through the tensor-core kernel (rank-local flat16 / norm planes)
    bank1, labels1 = synthetic_bank(2000, 1, 28, nlabels=10, seed=1)
    x1 = torch.randn(2, 1, 28, 28, generator=torch.Generator().manual_seed(5)).to(dev)
    kw = dict(kernel_size=5, batch_size=2000, image_size=28, schedule=cosine_noise_schedule, precision="auto")
    full = LocalScoreModule((bank1, labels1), **kw)
    shard = LocalScoreModule((bank1, labels1), process_group=dist.group.WORLD, **kw)
    assert shard.engine(dev).ls_umma_supported(5, 1)
    for label in (None, torch.tensor([3])):
        s0 = full(torch.full((2,), 0.5), x1, label=label, device=dev)
        s1 = shard(torch.full((2,), 0.5), x1, label=label, device=dev)
        err = float((s0 - s1).abs().max())
        worst = max(worst, err)
        assert err < 1e-3, ("LS C=1", label, err)
    # a class with a single image: with two ranks one shard is empty and must contribute the neutral element
    few_labels = labels.clone()
    few_labels[few_labels == 9] = 0
    few_labels[7] = 9
    for cls in (LocalEquivScoreModule, LocalEquivBordersScoreModule):
        full = cls((bank, few_labels), kernel_size=5, batch_size=64, schedule=cosine_noise_schedule)
        shard = cls((bank, few_labels), kernel_size=5, batch_size=64, schedule=cosine_noise_schedule,
                    process_group=dist.group.WORLD)
        s0 = full(torch.full((2,), 0.5), x, label=torch.tensor([9]), device=dev)
        s1 = shard(torch.full((2,), 0.5), x, label=torch.tensor([9]), device=dev)
        assert float((s0 - s1).abs().max()) < 1e-3, "empty shard"
    scales = [3, 3, 3, 5, 7, 9, 13, 17]
    full = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule)
    shard = LocalEquivScoreModule((bank, labels), kernel_size=3, batch_size=64, schedule=cosine_noise_schedule,
                                  process_group=dist.group.WORLD)
    m0, m1 = ScheduledScoreMachine(full, scales=scales), ScheduledScoreMachine(shard, scales=scales)
    o0 = m0(x, label=torch.tensor([2]), device=dev)
    o1 = m1(x, label=torch.tensor([2]), device=dev)
    o1b = m1(x, label=torch.tensor([2]), device=dev)          # graph replay with the captured all-gathers
    assert torch.equal(o1, o1b)
    # every rank holds only the images it owns
    nloc = shard.bank.N_local
    assert abs(nloc - 3000 / world) <= 10, nloc
    err = float((o0 - o1).abs().max())
    assert err < 1e-3, err
    gathered = [torch.empty_like(o1) for _ in range(world)]
    dist.all_gather(gathered, o1)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "x diverged across ranks"
    if rank == 0:
        print(f"dist check ok: world={world} worst score diff {worst:.2e}, trajectory diff {err:.2e}, ranks bit-identical, "
              f"{nloc} of 3000 images resident per rank")
    # orderly teardown: graphs (captured NCCL nodes) first, then the process group -- no os._exit
    from convolutional_diffusion_b200.distributed import shutdown
    m0.release_graphs()
    m1.release_graphs()
    shutdown()
    if rank == 0:
        print("teardown ok")


if __name__ == "__main__":
    main()
