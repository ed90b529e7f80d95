"""Per-step kernel sizes shipped with the reference as pickled Python lists
(`/root/reference/checkpoints/scales_*.pt`, loaded at `scripts/els_script.py:119-127`, read as `scales[i]` at
t = i/nsteps by `ScheduledScoreMachine.forward`, `src/utils/idealscore.py:95`; index 0 is never read).
Reproduced here as data so drivers work without the checkpoint files; `load_scales` also accepts a path to a
`.pt` file in the reference's format."""

SCALES = {
    "CIFAR10_ResNet_circular_conditional": [3, 3, 3, 3, 5, 5, 5, 5, 7, 7, 7, 9, 9, 11, 11, 13, 13, 15, 15, 15],
    "CIFAR10_ResNet_zeros_conditional": [3, 3, 3, 3, 5, 5, 5, 7, 7, 7, 7, 9, 9, 11, 11, 13, 15, 17, 17, 17],
    "CIFAR10_UNet_zeros_conditional": [3, 3, 3, 3, 5, 5, 5, 5, 7, 7, 7, 7, 9, 9, 11, 13, 13, 15, 17, 25],
    "CelebA_ResNet_zeros": [3, 3, 3, 5, 5, 5, 5, 5, 7, 7, 9, 9, 9, 11, 11, 11, 13, 15, 19, 19],
    "CelebA_UNet_zeros": [3, 3, 3, 3, 3, 3, 3, 5, 5, 5, 5, 5, 7, 7, 9, 9, 9, 13, 19, 27],
    "FashionMNIST_ResNet_zeros_conditonal": [3, 3, 5, 5, 7, 7, 9, 9, 9, 11, 11, 11, 13, 13, 13, 15, 15, 15, 17, 17],
    "FashionMNIST_UNet_zeros_conditonal": [3, 3, 5, 5, 7, 7, 7, 9, 11, 11, 11, 13, 15, 15, 21, 23, 25, 25, 25, 25],
    "MNIST_ResNet_circular": [3, 3, 5, 5, 5, 7, 7, 9, 9, 9, 11, 11, 11, 11, 9, 9, 7, 7, 3, 3],
    "MNIST_ResNet_zeros": [3, 3, 5, 5, 5, 7, 7, 7, 9, 9, 11, 11, 11, 11, 13, 15, 15, 15, 15, 15],
    "MNIST_UNet_zeros": [3, 3, 3, 5, 5, 7, 7, 7, 9, 9, 9, 11, 13, 15, 17, 21, 23, 23, 25, 27],
}


def load_scales(name_or_path):
    """`CIFAR10_ResNet_zeros_conditional` style name, or a path to a reference-format `scales_*.pt` file."""
    import os
    if name_or_path in SCALES:
        return list(SCALES[name_or_path])
    if os.path.isfile(name_or_path):
        import torch
        return [int(v) for v in torch.load(name_or_path, weights_only=False)]
    raise KeyError(f"unknown scales {name_or_path!r}; known: {sorted(SCALES)}")
