"""Per-step cost of sample_edm beyond its two network evaluations: python scripts/sampler_overhead.py [B] [steps]
(CUDA events; run under `ncu --metrics gpu__time_duration.sum` to list the non-network kernels of a step)."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.config import compose  # noqa: E402
from mcedm_b200.mcedm import PlMcedm  # noqa: E402
from mcedm_b200.utils import randomize_zero_init  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(pl.model, 2)
pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
pl = pl.to(dev).eval()
sp = copy.deepcopy(cfg.diff_sampler)
sp.timesteps = steps
cond = torch.randn(B, 2, 128, 128, device=dev)
mask = torch.zeros(B, 2, 128, 128, device=dev)
mask[:, 1] = 1.0
hu = torch.zeros(B, 2, 128, 128, device=dev)
for _ in range(2):
    pl.sample_edm(hu, cond, mask, sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pl.sample_edm(hu, cond, mask, sp)
e1.record()
torch.cuda.synchronize()
t_traj = e0.elapsed_time(e1)
unet = pl.ema_model.ma_model
eng = unet.engine()
x = torch.randn(B, 2, 128, 128, device=dev)
nl = torch.tensor([0.1], device=dev)
out = torch.empty(B, 2, 128, 128, device=dev)
xc, cc = eng._check_inputs(x, nl, cond)[0::2]
eng.forward_static(xc, nl, cc, out)
e0.record()
for _ in range(20):
    eng.forward_static(xc, nl, cc, out)
e1.record()
torch.cuda.synchronize()
t_eval = e0.elapsed_time(e1) / 20
n_eval = 2 * steps - 1
print(f"B={B}: trajectory of {steps} steps {t_traj:.2f} ms; evaluation {t_eval:.3f} ms x {n_eval} = {t_eval * n_eval:.2f} ms; "
      f"outside the network: {(t_traj - t_eval * n_eval) / steps:.3f} ms per step "
      f"({100 * (t_traj - t_eval * n_eval) / t_traj:.1f} % of the trajectory)")

# ---- components of one step outside the network, timed alone (20 launches each)
from mcedm_b200 import _lib as L  # noqa: E402

lib = L.lib()
f64 = dict(device=dev, dtype=torch.float64)
xa, xb, xc2, xd = (torch.randn(B, 2, 128, 128, **f64) for _ in range(4))
Fb, xin = torch.randn(B, 2, 128, 128, device=dev), torch.empty(B, 2, 128, 128, device=dev)
tot = xa.numel()
st = L.stream_ptr()


def timed(fn, n=20):
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


parts = {
    "randn fp64 (torch.randn_like, :608)": lambda: torch.randn_like(xa),
    "edm_churn": lambda: L.check(lib.mcedm_edm_churn(L.ptr(xa), L.ptr(xb), L.ptr(mask), 1.0, 0.5, tot, L.ptr(xc2), L.ptr(xin), st)),
    "edm_euler": lambda: L.check(lib.mcedm_edm_euler(L.ptr(xa), L.ptr(Fb), L.ptr(mask), 2.0, 1.5, 0.2, 0.9, 0.5, tot, L.ptr(xd), L.ptr(xc2), L.ptr(xin), None, st)),
    "edm_correct": lambda: L.check(lib.mcedm_edm_correct(L.ptr(xa), L.ptr(xc2), L.ptr(Fb), L.ptr(xd), L.ptr(mask), 2.0, 1.5, 0.2, 0.9, tot, L.ptr(xb), None, st)),
    "noise-label copy": lambda: nl.copy_(torch.zeros(1, device=dev)),
}
print("  ".join(f"{k}: {timed(v):.1f} us" for k, v in parts.items()))
