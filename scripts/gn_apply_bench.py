"""Times mcedm_gn_apply16 in the five shapes the fused evaluation uses: python scripts/gn_apply_bench.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcedm_b200 import _lib as L
lib = L.lib(); dev = torch.device("cuda:0"); dt = torch.float16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def geom(H, W):
    if W > 64:
        return 0, 0
    p, b = C.c_int(0), C.c_int(0)
    L.check(lib.mcedm_flat_geometry(H, W, C.byref(p), C.byref(b)))
    return p.value, b.value


def buf(H, W, flat=True):
    p, b = geom(H, W) if flat else (0, 0)
    n = B * b if p else B * H * W
    return torch.randn(n, 64, device=dev).to(dt), p, b


def run(name, Hin, Win, rs, pooled, dense_out=False):
    x, ip, ib = buf(Hin, Win)
    Ho, Wo = (Hin * 2, Win * 2) if rs == 1 else (Hin // 2, Win // 2) if rs == 2 else (Hin, Win)
    o, op, ob = buf(Ho, Wo, flat=not dense_out)
    pl = buf(Ho, Wo)[0] if pooled else None
    coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
    f = lambda: L.check(lib.mcedm_gn_apply16(L.ptr(x), ip, ib, L.ptr(coef), 1, rs, B, Hin, Win, op, ob, L.ptr(o), L.ptr(pl), 1,
                                             L.stream_ptr()))
    for _ in range(3):
        f()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    by = (B * Hin * Win + B * Ho * Wo * (2 if pooled else 1)) * 128.0
    print(f"{name:34s} {us:7.1f} us  {by / us / 1e3:7.0f} GB/s  ({by / 1e6:.0f} MB)")


run("128x128 -> 64x64 down (+pooled raw)", 128, 128, 2, True)
run("64x64 -> 128x128 up", 64, 64, 1, False)
run("64x64 -> 32x32 down (+pooled raw)", 64, 64, 2, True)
run("32x32 -> 64x64 up", 32, 32, 1, False)
run("32x32 same, dense out (qkv input)", 32, 32, 0, False, dense_out=True)
