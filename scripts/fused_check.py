"""Fused (16-bit activations, GroupNorm inside the convs) vs unfused inference plan: agreement with each other and
with the CPU oracle at a small batch, then graph-replay timing of both at a large batch.
python scripts/fused_check.py [B_time]"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
from mcedm_b200.adm_blocks import DhariwalUNet
from mcedm_b200.config import compose
from mcedm_b200.utils import randomize_zero_init, rel_l2
from oracle import edm_oracle as O

dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(net, 2)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
net = net.to(dev).eval()
eng = net.engine()
g = torch.Generator().manual_seed(0)
B = 3
x = torch.randn(B, 2, 128, 128, generator=g)
c = torch.randn(B, 2, 128, 128, generator=g)
nl = torch.tensor([0.3, -1.0, 0.7])
with torch.no_grad():
    ref = O.unet_forward(sd, dict(cfg.model.hparams.model), x, nl, c)
    outs = {}
    for fused in (False, True):
        eng.fused = fused
        outs[fused] = net(x.to(dev), nl.to(dev), c.to(dev)).cpu()
        L.check_watchdog()
    eng.fused, eng.precision = True, "fp32"
    out32 = net(x.to(dev), nl.to(dev), c.to(dev)).cpu()
    eng.precision = "fp16"
    L.check_watchdog()
print(f"fp32-accuracy plan vs oracle {rel_l2(out32, ref):.3e}")
print(f"unfused vs oracle {rel_l2(outs[False], ref):.3e}   fused vs oracle {rel_l2(outs[True], ref):.3e}   "
      f"fused vs unfused {rel_l2(outs[True], outs[False]):.3e}", flush=True)
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = torch.randn(Bt, 2, 128, 128, device=dev)
c = torch.randn(Bt, 2, 128, 128, device=dev)
nl = torch.tensor([0.3], device=dev)
out = torch.empty(Bt, 2, 128, 128, device=dev)
for fused in (False, True, "fp32"):
    eng.fused = bool(fused)
    eng.precision = "fp32" if fused == "fp32" else "fp16"
    eng._fmt = eng.infer_fmt
    eng.forward_static(x, nl, c, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        eng.forward_static(x, nl, c, out)
    e1.record()
    torch.cuda.synchronize()
    L.check_watchdog()
    ms = e0.elapsed_time(e1) / 20
    print(f"fused={fused} B={Bt}: {ms:.3f} ms per evaluation = {18.797e9 * Bt / ms / 1e9:.1f} TFLOP/s, "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)
