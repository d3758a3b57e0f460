"""Attention overflow fallback at the bench's batch (B = 256): flagged tiles spread so that fallback CTAs process
a flagged tile in their 2nd / 3rd strided iteration and some CTAs process two: python scripts/attn_overflow_repro.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcedm_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0")
B, Lq = 256, 1024
g = torch.Generator().manual_seed(9)
q = torch.randn(B, Lq, 64, generator=g) * 1.5
k = torch.randn(B, Lq, 64, generator=g)
v = torch.randn(B, Lq, 64, generator=g)
hot = [0, 20, 38, 255]
for s in hot:
    k[s] = k[s] * 0.05
    k[s, 896:] = q[s, :128] * 1.2
qkv = torch.cat([q, k, v], dim=2).to(dev).half().contiguous()
out = torch.full((B, Lq, 64), float("nan"), device=dev, dtype=torch.float16)
for it in range(3):
    L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out), None, 1, L.stream_ptr()), "attention")
    torch.cuda.synchronize()
    L.check_watchdog()
    print("call", it, "ok", flush=True)
for s in hot + [1, 100]:
    qd, kd, vd = qkv[s].double().split(64, dim=1)
    ref = torch.softmax(qd @ kd.t() / 8.0, dim=1) @ vd
    err = float((out[s].double() - ref).norm() / ref.norm())
    print(f"sample {s}: rel L2 {err:.2e}", flush=True)
    assert err < 2e-3
print("PASS")
