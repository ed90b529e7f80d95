"""Import the real reference (build container only).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so
nothing under `-m gpu`, `smoke()` or `bench.py` may call this; it is used by
`oracle/make_golden.py` and by CPU tests that are skipped when the
reference is absent.

The reference imports matplotlib at module scope (idealscore.py:4-5), which
this image does not have; empty stand-in modules are injected first.
"""
from __future__ import annotations

import os
import sys
import types

REF_DIR = os.environ.get("REF_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "src", "utils", "idealscore.py"))


def load():
    """Returns the reference's `src.utils.idealscore` module."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_DIR}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return importlib.import_module("src.utils.idealscore")


def load_models():
    """Returns the reference's `src.models` (the DDIM wrapper whose sample() holds the DDPM branch, models.py:48-64)."""
    load()
    import importlib
    return importlib.import_module("src.models")


class TensorBank:
    """Minimal map-style dataset yielding (image [C,H,W] float32, int label), the protocol the
    reference's DataLoader consumes (idealscore.py:142,390,489)."""

    def __init__(self, images, labels):
        self.images, self.labels = images, labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        return self.images[i], int(self.labels[i])
