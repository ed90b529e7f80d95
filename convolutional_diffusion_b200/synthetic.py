"""Seeded synthetic stand-ins for the datasets the reference trains on.

The real datasets (`src/utils/data.py:9-122`: MNIST / CIFAR10 / FashionMNIST /
CelebA through torchvision with `download=True`) are unreachable offline, so
banks of the same *shape and value range* are generated instead: low-pass
filtered Gaussian noise, rescaled per image, quantised to the 8-bit grid and
normalised with mean 0.5 / std 0.5 exactly like `ToTensor()+Normalize(0.5,0.5)`
(`data.py:63-70`), so values lie in [-1, 1] on the k/127.5-1 lattice.
Labels are uniform in [0, nlabels).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SHAPES = {
    # name: (channels, image_size, n_train, nlabels)    reference metadata: data.py:59-122
    "mnist": (1, 32, 60000, 10),        # data.py:66 resizes MNIST to 32x32
    "mnist28": (1, 28, 60000, 10),      # BASELINE.json cfg-1/2 quote the native 28x28 shape
    "fashionmnist": (1, 32, 60000, 10),
    "cifar10": (3, 32, 50000, 10),
    "celeba": (3, 32, 50000, 1),
    "celeba64": (3, 64, 50000, 1),
}


def synthetic_bank(n, channels, size, nlabels=10, seed=0, smooth=3, quantize8=True, chunk=8192):
    """Returns (images [n,C,H,W] float32 in [-1,1], labels [n] int64), deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(n, channels, size, size, dtype=torch.float32)
    box = torch.ones(1, 1, 3, 3) / 9.0
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        z = torch.randn(m * channels, 1, size, size, generator=g)
        for _ in range(smooth):
            z = F.conv2d(F.pad(z, (1, 1, 1, 1), mode="reflect"), box)
        z = z.view(m, channels * size * size)
        lo, hi = z.min(1, keepdim=True).values, z.max(1, keepdim=True).values
        z = (z - lo) / (hi - lo).clamp_min(1e-12)                      # [0,1]
        if quantize8:
            z = torch.round(z * 255.0) / 255.0                         # ToTensor() of a uint8 image
        out[s:s + m] = ((z - 0.5) / 0.5).view(m, channels, size, size)  # Normalize(0.5, 0.5)
    labels = torch.randint(0, max(nlabels, 1), (n,), generator=g)
    return out, labels


def synthetic_dataset(name, n=None, seed=0):
    c, size, n_full, nlabels = SHAPES[name]
    images, labels = synthetic_bank(n_full if n is None else n, c, size, nlabels, seed)
    return images, labels, dict(name=name, num_channels=c, image_size=size, nlabels=nlabels)


def noisy_query(images, beta, batch, seed):
    """x = a*T_j + sqrt(beta)*eps for random bank images j: a query with realistic softmax entropy
    at noise level beta."""
    g = torch.Generator().manual_seed(seed)
    j = torch.randint(0, images.shape[0], (batch,), generator=g)
    eps = torch.randn(batch, *images.shape[1:], generator=g)
    return (1.0 - beta) ** 0.5 * images[j] + beta ** 0.5 * eps


class TensorDataset2:
    """Map-style dataset yielding (image, int label): the protocol the score modules accept
    (`idealscore.py:142,390,489` wrap it in a DataLoader)."""

    def __init__(self, images, labels):
        self.images, self.labels = images, labels

    def __len__(self):
        return self.images.shape[0]

    def __getitem__(self, i):
        return self.images[i], int(self.labels[i])
