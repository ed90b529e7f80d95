"""Host-side profile of one bbELS module call at the cfg-4 shape (where does the time outside the kernels go)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalEquivBordersScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402

dev = torch.device("cuda", 0)
bank, labels = synthetic_bank(50000, 3, 32, nlabels=10, seed=0)
mod = LocalEquivBordersScoreModule((bank, labels), kernel_size=17, batch_size=64, image_size=32, schedule=cosine_noise_schedule)
x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(1)).to(dev)
t = torch.tensor([0.9])
for _ in range(2):
    mod(t, x, device=dev)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    mod(t, x, device=dev)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        mod(t, x, device=dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
