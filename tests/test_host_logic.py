"""CPU tests: host-side selection logic against the oracle's replay of the DataLoader loop, the C-ABI
library's exported symbols, the DDIM coefficient table."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from oracle import score_oracle as so
from convolutional_diffusion_b200 import selection as sel
from convolutional_diffusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind", ["LS", "ELS", "bbELS"])
@pytest.mark.parametrize("n,bs,label,ms", [
    (40, 16, None, None), (40, 16, 3, None), (48, 12, 1, 30), (48, 12, None, 30), (100, 7, 2, 50),
    (100, 100, None, None), (33, 8, 0, 8), (64, 16, 9, None), (64, 16, None, 0), (10, 64, 1, 5),
])
def test_select_matches_oracle(kind, n, bs, label, ms):
    rng = np.random.default_rng(n * 31 + bs)
    labels = rng.integers(0, 4, n)
    for order in (None, rng.permutation(n)):
        i0, w0 = so.select_bank(kind, labels, label, bs, ms, order)
        i1, w1 = sel.select(kind, labels, label, bs, ms, order)
        assert np.array_equal(i0, i1)
        assert np.allclose(w0, w1)


def test_shard_partition():
    idx = np.arange(103)
    logw = -np.log(np.arange(1, 104, dtype=np.float64))
    seen = []
    for r in range(8):
        i, w = sel.shard(idx, logw, r, 8)
        assert np.allclose(w, logw[i])
        seen.append(i)
    assert np.array_equal(np.sort(np.concatenate(seen)), idx)
    assert max(len(s) for s in seen) - min(len(s) for s in seen) <= 1


def test_rank_ownership_is_a_balanced_partition_of_every_selection():
    """Bank sharding by image ownership (selection.assign_ranks): the owner depends on the image only, every selection
    -- conditional, unconditional, with max_samples, shuffled -- is partitioned by it, and the parts are balanced per class
    (sizes differ by at most one) so that no rank idles in a class-conditional run."""
    rng = np.random.default_rng(3)
    labels = rng.integers(0, 10, 5003)
    for world in (2, 4, 8):
        owner = sel.assign_ranks(labels, world)
        assert owner.min() == 0 and owner.max() == world - 1
        for c in range(10):
            counts = np.bincount(owner[labels == c], minlength=world)
            assert counts.max() - counts.min() <= 1, (world, c, counts)
        for kind, label, bs, ms, order in (("ELS", 4, 64, None, None), ("ELS", None, 64, 3000, None),
                                           ("LS", 7, 256, None, rng.permutation(5003)), ("bbELS", None, 64, None, None)):
            idx, logw = sel.select(kind, labels, label, bs, ms, order)
            parts = [sel.shard(idx, logw, r, world, owner) for r in range(world)]
            assert np.array_equal(np.sort(np.concatenate([p[0] for p in parts])), np.sort(idx))
            for pi, pw in parts:
                pos = {int(v): q for q, v in enumerate(idx)}
                assert np.allclose(pw, logw[[pos[int(v)] for v in pi]])        # log-weights travel with their images
            if ms is None and order is None:
                sizes = [len(p[0]) for p in parts]
                assert max(sizes) - min(sizes) <= (10 if label is None else 1), (kind, label, sizes)
    assert np.array_equal(sel.assign_ranks(labels, 1), np.zeros(5003, dtype=np.int64))

def test_shuffle_order_matches_dataloader():
    """LS hard-codes shuffle=True (idealscore.py:489): reproduce the DataLoader's permutation for a given
    global RNG state."""
    from torch.utils.data import DataLoader, TensorDataset
    ds = TensorDataset(torch.arange(50))
    torch.manual_seed(123)
    got = [int(v) for (b,) in DataLoader(ds, batch_size=7, shuffle=True) for v in b]
    torch.manual_seed(123)
    mine = sel.dataloader_shuffle_order(50)
    assert got == [int(v) for v in mine]


def test_ddim_coefficients_match_oracle():
    from convolutional_diffusion_b200.machine import ddim_coefficients
    for nsteps in (6, 20):
        mine = ddim_coefficients(nsteps)
        ref = so.machine_coeffs(nsteps)
        assert len(mine) == nsteps - 1
        for (i, bt, cx, cmu), (j, bt2, bp2, cx_eps, ce) in zip(mine, ref):
            assert i == j and abs(bt - bt2) < 1e-6
            # x' = cx_eps x + ce eps with eps = (x - a mu)/sqrt(bt)  ==>  coefficients of x and mu
            a = math.sqrt(1 - bt2)
            assert abs(cx - (cx_eps + ce / math.sqrt(bt2))) < 1e-5
            assert abs(cmu - (-ce * a / math.sqrt(bt2))) < 1e-5


def test_schedules_match_golden():
    from conftest import load_case, GOLDEN
    from convolutional_diffusion_b200 import cosine_noise_schedule, exponential_schedule
    c = load_case(os.path.join(GOLDEN, "schedule.npz"))
    t = torch.from_numpy(c["t"])
    assert np.allclose(cosine_noise_schedule(t).numpy(), c["cosine"], atol=1e-7)
    assert np.allclose(exponential_schedule(t).numpy(), c["exponential"], atol=1e-7)


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/cdscore.h declares."""
    from convolutional_diffusion_b200 import build
    build.build()
    header = open(os.path.join(ROOT, "include", "cdscore.h")).read()
    declared = set(re.findall(r"\b(cds_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.cds_abi_version() == 2
    # geometry query is host-only arithmetic: CIFAR shape, k=17, two query passes fits in 227 KB
    assert 0 < loaded.cds_els_umma_smem_bytes(3, 32, 32, 17, 2, 1) <= 227 * 1024
    assert 0 < loaded.cds_els_umma_smem_bytes(3, 64, 64, 17, 1, 1) <= 227 * 1024    # 64x64: staged in row bands
    assert loaded.cds_els_umma_smem_bytes(3, 128, 128, 5, 1, 1) == 0     # larger images fall back to the SIMT kernel
    assert loaded.cds_els_umma_smem_bytes(3, 32, 32, 4, 1, 1) == 0       # even kernel sizes are rejected


def test_edge_and_ls_tensor_core_geometries():
    """Host-only geometry queries of the tensor-core bbELS edge-band and LS kernels: what is supported fits 227 KB of shared
    memory, what is not reports 0 (the engine then keeps the exact SIMT kernels)."""
    lib = _lib.load()
    for H, C in ((32, 3), (28, 1), (64, 3)):
        for k in range(3, min(H, 28), 2):
            assert 0 < lib.cds_bbels_edge_umma_smem_bytes(C, H, H, k, 1) <= 227 * 1024, (H, C, k)
    assert 0 < lib.cds_bbels_edge_umma_smem_bytes(3, 32, 32, 17, 2) <= 227 * 1024      # cfg-4 with two query passes
    assert lib.cds_bbels_edge_umma_smem_bytes(3, 32, 32, 31, 2) == 0                      # query blocks too large: SIMT kernel
    assert lib.cds_bbels_edge_umma_smem_bytes(3, 32, 24, 7, 1) == 0                       # square images only
    assert lib.cds_bbels_edge_umma_smem_bytes(3, 32, 32, 32, 1) == 0                      # k >= H is the LS delegate
    assert lib.cds_edge_plane_halves(10, 3, 32) == 10 * 4 * 3 * 4 * 32 * 8
    assert lib.cds_edge_norms_halves(10, 32, 17) == 10 * 4 * 8 * 16 * 8
    for k in (3, 5, 9, 17, 27):
        assert 0 < lib.cds_ls_umma_smem_bytes(1, 28, 28, k, 1) <= 227 * 1024, k            # cfg-1 shapes, one query pass
    assert 0 < lib.cds_ls_umma_smem_bytes(1, 28, 28, 5, 2) <= 227 * 1024
    assert lib.cds_ls_umma_smem_bytes(1, 28, 28, 17, 2) == 0      # two planes of the query operand exceed TMEM: SIMT kernel
    assert lib.cds_ls_umma_smem_bytes(3, 32, 32, 5, 1) == 0       # multi-channel LS (bbELS corners) stays on the SIMT kernel
    assert lib.cds_ls_umma_smem_bytes(1, 28, 28, 55, 1) == 0      # whole-image window (IS)
    assert lib.cds_ls_plane_elems(7, 28, 28) == 7 * 792 and lib.cds_ls_norms_elems(7, 28, 28) == 7 * 896


def _umma_geometries(ks, variant, passes=1):
    """Runs the geometry selection of cds_els_partials_umma (printed with CDS_DEBUG_GEOM) in a subprocess without a
    GPU: the launch itself fails, the pointers are dummies."""
    import subprocess
    import sys
    code = (
        "import ctypes, sys\n"
        "sys.path.insert(0, %r)\n"
        "from convolutional_diffusion_b200 import _lib\n"
        "lib = _lib.load()\n"
        "one = ctypes.c_void_p(16)\n"
        "for k in %r:\n"
        "    lib.cds_els_partials_umma(1, None, 4, 3, 32, 32, k, None, None, None, one, 255.0, None, None, None, 100, 9, %d,\n"
        "                              %d, None, None, None, None, None)\n" % (ROOT, tuple(ks), passes, variant))
    env = dict(os.environ, CDS_DEBUG_GEOM="1", CUDA_VISIBLE_DEVICES="")
    env.pop("CDS_ELS_MIXED", None)
    env.pop("CDS_PV_MAX_K", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    geo = {}
    for line in r.stderr.splitlines():
        m = re.search(r"els_umma k=(\d+) .* pv=(\d) mixed=(\d) G=(\d+) chunks=(\d+) nvb=(\d+) n_mma=(\d+) n_tmem=(\d+) stages=(\d+) smem=(\d+)", line)
        if m:
            v = [int(g) for g in m.groups()]
            geo[v[0]] = dict(pv=v[1], mixed=v[2], G=v[3], chunks=v[4], nvb=v[5], n_mma=v[6], n_tmem=v[7], stages=v[8], smem=v[9])
    assert set(geo) == set(ks), (r.stderr[-500:], r.stdout[-200:])
    return geo


def test_umma_geometry_choices():
    """Host-side tiling decisions of the tensor-core kernel, FMA-pipe epilogue: CIFAR shape, one query pass.  The mixed K
    layout must keep the band height of the vertical layout and is taken for k = 9, 11, 13, 17 (17/28, 26/34, 35/40,
    57/77 UMMAs per tile) but not for k = 15 (44/46)."""
    ks = (5, 9, 11, 13, 15, 17)
    geo = _umma_geometries(ks, variant=1)
    assert [geo[k]["pv"] for k in ks] == [0] * 6
    assert [geo[k]["mixed"] for k in ks] == [0, 1, 1, 1, 0, 1]
    assert [geo[k]["n_mma"] for k in ks] == [8, 17, 26, 35, 46, 57]
    assert [geo[k]["G"] for k in ks] == [28, 24, 22, 20, 18, 16]       # one band = all patch rows
    for k, g in geo.items():
        assert g["stages"] == 2 and g["chunks"] == 1 and g["smem"] <= 227 * 1024, (k, g)
        assert 16 * g["G"] + 8 * g["n_tmem"] <= 512, (k, g)                              # TMEM columns


def test_umma_pv_geometry_choices():
    """The same for the P.V epilogue (both query-pass counts): same bands and UMMA counts as the FMA epilogue, two stages
    (the stage of a band is released one tile late), the 16-column O accumulator fits next to the two S buffers; "auto"
    takes P.V up to k = 9 and the FMA epilogue above."""
    ks = (3, 5, 7, 9, 11, 13, 15, 17)
    for passes in (1, 2):
        geo = _umma_geometries(ks, variant=2, passes=passes)
        ref = _umma_geometries(ks, variant=1, passes=passes)
        for k in ks:
            g = geo[k]
            assert g["pv"] == 1 and g["stages"] == 2 and g["smem"] <= 227 * 1024, (k, passes, g)
            assert (g["G"], g["chunks"], g["nvb"], g["n_mma"], g["mixed"]) == \
                (ref[k]["G"], ref[k]["chunks"], ref[k]["nvb"], ref[k]["n_mma"], ref[k]["mixed"]), (k, passes, g, ref[k])
            assert 16 * g["G"] + 16 + 8 * g["n_tmem"] <= 512, (k, g)
    auto = _umma_geometries(ks, variant=0)
    assert [auto[k]["pv"] for k in ks] == [1, 1, 1, 1, 0, 0, 0, 0]


def test_label_groups():
    """Per-sample labels: grouping helper behind modules.forward / ScheduledScoreMachine.forward."""
    from convolutional_diffusion_b200.modules import _label_groups
    assert _label_groups(None, 4) is None
    assert _label_groups(torch.tensor([3]), 4) is None                  # one label for the whole call (the reference)
    assert _label_groups(torch.tensor([3, 3, 3]), 3) is None            # all equal: a single evaluation
    assert _label_groups(torch.tensor([2, 0, 2, 1]), 4) == {2: [0, 2], 0: [1], 1: [3]}
    assert _label_groups([1, 0], 2) == {1: [0], 0: [1]}
    assert _label_groups(5, 4) is None
    with pytest.raises(ValueError):
        _label_groups(torch.tensor([0, 1, 2]), 4)


def test_no_cpu_fallback():
    """Without CUDA the product path must fail loudly, not fall back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from convolutional_diffusion_b200 import LocalEquivScoreModule
    mod = LocalEquivScoreModule((torch.zeros(4, 3, 8, 8), torch.zeros(4, dtype=torch.long)), kernel_size=3)
    with pytest.raises(RuntimeError):
        mod(torch.tensor([0.5]), torch.zeros(1, 3, 8, 8), device=torch.device("cpu"))
    with pytest.raises(RuntimeError):
        mod.engine("cuda")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "convolutional_diffusion_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
