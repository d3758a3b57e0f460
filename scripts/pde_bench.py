"""Times the PDE-residual kernels (K6) on a float64 NCHW state batch: python scripts/pde_bench.py [B] [n]
(CUDA events around n back-to-back launches; run under `ncu --metrics gpu__time_duration.sum` for pure kernel time)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.nn_misc import Normalizer  # noqa: E402
from mcedm_b200.pde_loss import DarcyLoss, SweFvLoss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
x = torch.randn(B, 2, 128, 128, device=dev, dtype=torch.float64) * 0.1 + 1.0
nh, nu = Normalizer(torch.tensor(1.5), torch.tensor(0.3)).to(dev), Normalizer(torch.tensor(0.0), torch.tensor(0.2)).to(dev)
f, d = SweFvLoss(Tn=0.128, x_min=-0.5, x_max=0.5), DarcyLoss()
cells = B * 128 * 128
for name, fn, byts in (("swe_fv_loss", lambda: f.residual(x[:, 0], x[:, 1], nh, nu), cells * 16.0),
                       ("swe_fv_loss+matrix", lambda: f.residual(x[:, 0], x[:, 1], nh, nu, want_matrix=True), cells * 24.0),
                       ("swe_fv_grad", lambda: f.gradient(x[:, 0], x[:, 1], nh, nu, mode=0), cells * 24.0),
                       ("swe_fv_grad mean", lambda: f.gradient(x[:, 0], x[:, 1], nh, nu, mode=1), cells * 20.0),
                       ("darcy_loss", lambda: d.residual(x[:, 0], x[:, 1], nh, nu), cells * 16.0)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:22s} B={B}: {ms * 1e3:8.1f} us per call  {byts / ms / 1e6:8.1f} GB/s (algorithmic bytes)")
