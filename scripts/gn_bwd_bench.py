"""Times mcedm_gn_bwd16 (16-bit training plan) per level: python scripts/gn_bwd_bench.py [B]
Reports us per call and GB/s over the algorithmic bytes (pass 1: x + dy; pass 2: x + dy + add0 + dx16)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
lib = L.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def geom(H, W):
    p, b = C.c_int(0), C.c_int(0)
    L.check(lib.mcedm_flat_geometry(H, W, C.byref(p), C.byref(b)))
    return p.value, b.value


for H in (128, 64, 32):
    W = H
    lay = geom(H, W) if W <= 64 else (0, 0)
    npos = B * lay[1] if W <= 64 else B * H * W
    x = torch.randn(npos, 64, device=dev).half()
    dy = torch.randn(npos, 64, device=dev).half()
    add0 = torch.randn(npos, 64, device=dev).half()
    dx16 = torch.zeros(npos, 64, device=dev, dtype=torch.float16)
    mr = torch.rand(B, 16, 2, device=dev) + 0.5
    gamma, beta = torch.randn(64, device=dev), torch.randn(64, device=dev)
    n_cta = lib.mcedm_gn_bwd16_ctas_per_img(H, W, B)
    red = torch.empty(B, n_cta, 64, 2, device=dev)
    kcoef = torch.empty(B, 192, device=dev)
    coefab = torch.randn(B, 128, device=dev)
    ticket = torch.zeros(3 * B, device=dev, dtype=torch.int32)
    dgb = torch.empty(B, 64, 2, device=dev)
    cs = torch.empty(B * n_cta, 64, device=dev)

    def call(with_add):
        L.check(lib.mcedm_gn_bwd16(L.ptr(dy), lay[0], lay[1], L.ptr(x), lay[0], lay[1], 1, L.ptr(mr), L.ptr(coefab), L.ptr(gamma),
                                   L.ptr(beta), None, 128, 64, 1, 0, B, H, W, L.ptr(red), L.ptr(kcoef), L.ptr(ticket),
                                   L.ptr(dgb), None, 128, L.ptr(add0) if with_add else None, 0, lay[0], lay[1], None, 1,
                                   None, L.ptr(dx16), None, L.ptr(cs), L.stream_ptr()))

    for with_add in (False, True):
        for _ in range(3):
            call(with_add)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call(with_add)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        us = ts[len(ts) // 2] * 1e3
        n = B * H * W * 64
        by = n * (4 + 6 + (2 if with_add else 0))
        print(f"{H:3d}x{W:<3d} B={B} add0={int(with_add)} ctas/img={n_cta:3d}: {us:7.1f} us per call (both passes), "
              f"{by / us / 1e3:6.0f} GB/s algorithmic")
