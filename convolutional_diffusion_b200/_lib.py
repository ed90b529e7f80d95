"""ctypes binding of the C ABI in include/cdscore.h (libcdscore.so, built in-tree by build.py).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CDS_LIB_PATH", os.path.join(_HERE, "libcdscore.so"))   # override: A/B builds only

KIND = {"LS": 0, "ELS": 1, "bbELS": 2}
PAD = {"zeros": 0, "circular": 1}
ELS_VARIANT = {"auto": 0, "fma": 1, "v2": 1, "pv": 2}      # CDS_ELS_AUTO / _FMA / _PV

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/cdscore.h one to one
SIGNATURES = {
    "cds_abi_version": (_i, []),
    "cds_last_error": (C.c_char_p, []),
    "cds_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "cds_pack_strip8": (_i, [_p, _i64, _i, _i, _i, _f, _i, _p, _p]),
    "cds_patch_norms": (_i, [_p, _i64, _i, _i, _i, _i, _p, _p]),
    "cds_pack_norm_plane": (_i, [_p, _i64, _i, _i, _i, _i, _p, _p]),
    "cds_partials_simt": (_i, [_i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "cds_score_vjp_simt": (_i, [_i, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "cds_ls_partials": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p]),
    "cds_ls_rows_supported": (_i, [_i, _i, _i, _i]),
    "cds_ls_rows_partials": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p]),
    "cds_ls_umma_smem_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "cds_ls_plane_elems": (_i64, [_i64, _i, _i]),
    "cds_ls_norms_elems": (_i64, [_i64, _i, _i]),
    "cds_pack_flat16": (_i, [_p, _i64, _i, _i, _i, _f, _p, _p]),
    "cds_pack_ls_norms": (_i, [_p, _i64, _i, _i, _i, _i, _p, _p]),
    "cds_ls_partials_umma": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "cds_bbels_edge_supported": (_i, [_i, _i, _i, _i]),
    "cds_bbels_edge_partials": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p]),
    "cds_bbels_edge_umma_smem_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "cds_edge_plane_halves": (_i64, [_i64, _i, _i]),
    "cds_edge_norms_halves": (_i64, [_i64, _i, _i]),
    "cds_pack_edge_plane": (_i, [_p, _i64, _i, _i, _f, _p, _p]),
    "cds_pack_edge_norms": (_i, [_p, _i64, _i, _i, _i, _p, _p]),
    "cds_bbels_edge_partials_umma": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "cds_els_partials_umma": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _i64, _i, _i, _i,
                                   _p, _p, _p, _p, _p]),
    "cds_els_partials_umma_window": (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _i64, _i, _i, _i,
                                          _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "cds_els_umma_smem_bytes": (_i64, [_i, _i, _i, _i, _i, _i]),
    "cds_els_umma_pv_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "cds_combine": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "cds_combine_packed": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "cds_finalize": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "cds_ddim_step": (_i, [_p, _p, _p, _p, _i, _i64, _p]),
    "cds_finish": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "cds_randn_philox": (_i, [_p, _i64, _p, _i, _p]),
}

_lib = None


def load():
    """Returns the loaded library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m convolutional_diffusion_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cds_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
