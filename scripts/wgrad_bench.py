"""conv_wgrad16_fused at 128x128 (dense): time vs batch and with / without the in-kernel transform:
python scripts/wgrad_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcedm_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0"); dt = torch.float16


def run(B, H, W, xf, taps=9, n=20):
    dy = (torch.randn(B, H, W, 64, device=dev) * 0.1).to(dt)
    a = torch.randn(B, H, W, 64, device=dev).to(dt)
    coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
    nc = lib.mcedm_wgrad_ctas(B, H, W)
    part = torch.empty(nc * taps * 4096, device=dev)
    f = lambda: L.check(lib.mcedm_conv_wgrad16_fused(L.ptr(dy), 0, 64, 0, L.ptr(a), 0, 64, 0, L.ptr(coef) if xf else None, 1, B, H, W,
                                                     taps, L.ptr(part), 1, L.stream_ptr()))
    for _ in range(3):
        f()
    torch.cuda.synchronize(); L.check_watchdog()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * H * W * 64 * 64 * taps
    print(f"B={B:4d} {H}x{W} taps={taps} xf={int(xf)} ctas={nc}: {us:7.1f} us  {fl / us / 1e6:6.0f} TFLOP/s  "
          f"({us * 1e-6 * 1.7e9 / (B * H / nc):.0f} cycles per row per CTA at 1.7 GHz)")


for B in (32, 128):
    for xf in (True, False):
        run(B, 128, 128, xf)
run(32, 128, 128, True, taps=1)
print("--- small problems (fixed cost)")
run(2, 128, 128, True)
run(2, 128, 128, False)
run(32, 64, 64, True)
run(32, 64, 64, False)
run(32, 32, 32, True)
run(32, 32, 32, False)
run(32, 32, 32, False, taps=1)
