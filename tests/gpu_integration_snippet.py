"""Executes the raw C-ABI binding shown in INTEGRATION.md section B (ctypes only, no package imports on the call path)
and compares it with the float64 oracle (not a pytest file).  Usage: python tests/gpu_integration_snippet.py"""
import ctypes
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from convolutional_diffusion_b200.synthetic import synthetic_bank  # noqa: E402  (test data only)
from oracle import score_oracle as so  # noqa: E402


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "convolutional_diffusion_b200", "libcdscore.so"))
    P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.cds_pack_strip8.argtypes = [P, L, I, I, I, F, I, P, P]
    lib.cds_pack_norm_plane.argtypes = [P, L, I, I, I, I, P, P]
    lib.cds_els_partials_umma.argtypes = [I, P, I, I, I, I, I, P, P, P, P, F, P, P, P, L, I, I, I, P, P, P, P, P]
    lib.cds_combine.argtypes = [P, P, P, I, I, I, I, P, P, P, P]
    lib.cds_finalize.argtypes = [P, P, P, P, P, I, I, I, I, I, I, P, P, P]
    lib.cds_last_error.restype = ctypes.c_char_p

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def ok(rc):
        assert rc == 0, lib.cds_last_error()

    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    bank, labels = synthetic_bank(48, 3, 32, nlabels=3, seed=1)
    images = bank.cuda().float().contiguous()
    N, C, H, W = images.shape
    for k, t in ((5, 0.2), (11, 0.6), (17, 0.9)):
        strip = torch.empty(N * C * H * W * 8, dtype=torch.float16, device="cuda")
        ok(lib.cds_pack_strip8(ptr(images), N, C, H, W, 255.0, 0, ptr(strip), stream))
        rows = torch.empty_like(strip)
        ok(lib.cds_pack_strip8(ptr(images), N, C, H, W, 255.0, 2, ptr(rows), stream))
        pn = torch.empty(N * H * W * 8, dtype=torch.float16, device="cuda")
        ok(lib.cds_pack_norm_plane(ptr(images), N, C, H, W, k, ptr(pn), stream))
        label, bs = 1, 16
        idx_np, logw_np = so.select_bank("ELS", labels.numpy(), label, bs, None)      # DataLoader bookkeeping
        idx = torch.as_tensor(idx_np, dtype=torch.int32).cuda()
        logw = torch.as_tensor(logw_np, dtype=torch.float32).cuda()
        n_sel = idx.numel()
        B = 2
        bt = float(so.cosine_beta(t))
        x = (math.sqrt(1 - bt) * bank[:B] + math.sqrt(bt) * torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(k))).cuda()
        beta = torch.full((B,), bt, device="cuda")
        S = 4
        m = torch.zeros(S, B, H * W, device="cuda")
        l = torch.zeros_like(m)
        acc = torch.zeros(S, B, C, H * W, device="cuda")
        ok(lib.cds_els_partials_umma(1, ptr(x), B, C, H, W, k, ptr(beta), ptr(strip), None, ptr(rows), 255.0, ptr(pn),
                                     ptr(idx), ptr(logw), n_sel, S, 2, 0, ptr(m), ptr(l), ptr(acc), None, stream))
        ok(lib.cds_combine(ptr(m), ptr(l), ptr(acc), S, B, C, H * W, ptr(m), ptr(l), ptr(acc), stream))
        score = torch.empty_like(x)
        ok(lib.cds_finalize(ptr(x), ptr(beta), ptr(m), ptr(l), ptr(acc), B, C, H, W, 0, 0, None, ptr(score), stream))
        torch.cuda.synchronize()
        for b in range(B):
            s_ref, _ = so.score("ELS", x[b].cpu().numpy(), bank.numpy()[idx_np], bt, k, logw_np)
            err = float(np.max(np.abs(score[b].cpu().double().numpy() - s_ref))) * bt / math.sqrt(1 - bt)
            assert err < 1e-3, (k, b, err)
        print(f"raw C ABI, k={k}: ok")


if __name__ == "__main__":
    main()
