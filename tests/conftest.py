import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def load_case(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def have_cuda():
    import torch
    return torch.cuda.is_available()


def replay_ddpm_noise(seed, shape, nsteps):
    """The Gaussian draws of `DDIM.sample(ddpm=True)` (src/models.py:62) for a golden generated from torch.manual_seed(seed):
    every step first opens a DataLoader iterator inside the score module (one int64 draw from the global RNG for its base
    seed, even with shuffle=False), then calls randn_like(x)."""
    import torch
    torch.manual_seed(int(seed))
    out = []
    for _ in range(nsteps):
        torch.empty((), dtype=torch.int64).random_()
        out.append(torch.randn(tuple(shape)))
    return out
