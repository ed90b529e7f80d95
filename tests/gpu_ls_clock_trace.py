"""Profile build only (CDS_LIB_PATH=.../libcdscore_prof.so CDS_LS_DEBUG=8): timeline of the first tiles of one CTA of the
tensor-core LS kernel."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convolutional_diffusion_b200 import LocalScoreModule, cosine_noise_schedule, _lib  # noqa: E402
from convolutional_diffusion_b200.synthetic import synthetic_bank, noisy_query  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("CDS_B", "10"))
bank, labels = synthetic_bank(60000, 1, 28, nlabels=10, seed=0)
mod = LocalScoreModule((bank, labels), kernel_size=5, batch_size=60000, image_size=28, schedule=cosine_noise_schedule, precision="auto")
eng = mod.engine(dev)
sel = mod.selection(None)
beta = cosine_noise_schedule(torch.tensor([0.8] * B)).to(dev, torch.float32)
x = noisy_query(bank[:64], float(beta[0]), B, seed=1).to(dev)
for _ in range(3):
    eng.ls_partials(x, beta, 5, sel, passes=1)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 64))()
lib = _lib.load()
lib.cds_debug_ls_clocks.argtypes = [ctypes.c_void_p]
assert lib.cds_debug_ls_clocks(buf) == 0
rows = [[buf[r * 64 + t] for t in range(64)] for r in range(8)]
t0 = min(v for r in rows for v in r if v > 0)
names = ["landed", "published", "buf free", "issued", "mma start", "mma issued", "epi start", "epi end"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for t in range(8, 40):
    print(f"{t:4d} " + " ".join(f"{rows[r][t] - t0:10d}" for r in range(8)))
