"""In-tree build of libcdscore.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m convolutional_diffusion_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcdscore.so")
SOURCES = ["api.cu", "simt_kernels.cu", "ls_kernel.cu", "ls_rows_kernel.cu", "bbels_edge.cu", "bbels_edge_umma.cu", "ls_umma.cu", "els_umma.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "--threads", "0"]      # --threads 0: compile the sources in parallel


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "cdscore.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, profile=False):
    """profile=True builds libcdscore_prof.so with -DCDS_PROFILE_SWITCHES (A/B switches and skip counters of the
    tcgen05 kernel; selected at run time with CDS_LIB_PATH) -- never the product library."""
    out = OUT.replace(".so", "_prof.so") if profile else OUT
    if not force and not profile and not _stale():
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    extra = os.environ.get("CDS_EXTRA_NVCC_FLAGS", "").split()       # A/B builds: -D switches of the kernels
    out = os.environ.get("CDS_BUILD_OUT", out)
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DCDS_PROFILE_SWITCHES"] if profile else []) \
        + extra + ["-o", out] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv))
