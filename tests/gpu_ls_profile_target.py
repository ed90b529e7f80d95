"""Profiling target: LS bank-streaming kernel, MNIST shape 60k x 1 x 28 x 28, k=5, B=1 (BASELINE config 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalScoreModule, cosine_noise_schedule  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402

bank, labels = synthetic_bank(60000, 1, 28, seed=0)
mod = LocalScoreModule((bank, labels), kernel_size=5, batch_size=60000, schedule=cosine_noise_schedule)
eng = mod.engine("cuda")
sel = mod.selection(None)
x = noisy_query(bank, 0.3, 1, seed=1).cuda()
beta = torch.full((1,), 0.3, device="cuda")
for _ in range(3):
    eng.ls_partials(x, beta, 5, sel)
torch.cuda.synchronize()
print("ok")
