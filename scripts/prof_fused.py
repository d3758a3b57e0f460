"""CUDA-event timing of the GroupNorm-fused conv kernels standalone: python scripts/prof_fused.py [B] [reps]
(also the ncu target for these kernels)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
from mcedm_b200.engine import pack_conv3x3

dev = torch.device("cuda:0")
lib = L.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dt = torch.float16


def timeit(fn, flops, bytes_):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    L.check_watchdog()
    ms = e0.elapsed_time(e1) / reps
    return f"{ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s  {bytes_ / ms / 1e6:7.1f} GB/s"


def rows(H, n_halo, n_ctr, N, res_mode, xf):
    n_total = 64 if N == 32 else N
    halos = [torch.randn(B, H, 128, 64, device=dev).to(dt) for _ in range(n_halo)]
    ctrs = [torch.randn(B, H, 128, 64, device=dev).to(dt) for _ in range(n_ctr)]
    coefs = [torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
             for _ in range(n_halo)]
    w = pack_conv3x3(torch.randn(n_total, 64 * n_halo, 3, 3, device=dev) / 24, dtype=dt)
    if n_ctr:
        w = torch.cat([w, pack_conv3x3(torch.randn(n_total, 64 * n_ctr, 1, 1, device=dev) / 8, dtype=dt)], 0).contiguous()
    bias = torch.randn(n_total, device=dev)
    res = torch.randn(B, H, 128, n_total, device=dev).to(dt) if res_mode == 1 else None
    out = torch.empty(B, H, 128, n_total, device=dev, dtype=dt)
    st = torch.empty(B * H, 4, n_total // 4, 2, device=dev)
    cptr = (C.c_void_p * n_halo)(*[c.data_ptr() for c in coefs]) if xf else None

    def fn():
        for n_off in range(0, n_total, N):
            L.check(lib.mcedm_conv_rows_fused(L.ptr_array(halos), cptr, n_halo, L.ptr_array(ctrs) if ctrs else None, n_ctr,
                                              L.ptr(w), L.ptr(bias), B, H, N, n_off, n_total, L.ptr(out), 1, L.ptr(res),
                                              res_mode, 0, 0, L.ptr(st), 1, L.stream_ptr()))
    px = B * H * 128
    flops = 2.0 * px * n_total * 64 * (9 * n_halo + n_ctr)
    bytes_ = px * 64 * 2.0 * (n_halo * (n_total // N) + n_ctr + (1 if res_mode else 0)) + px * n_total * 2.0
    print(f"rows  H={H} halo={n_halo} ctr={n_ctr} N={N} res={res_mode} xf={int(xf)}: " + timeit(fn, flops, bytes_), flush=True)


def flat(H, W, res_mode, xf, out_f32=0):
    pitch, blk = C.c_int(0), C.c_int(0)
    L.check(lib.mcedm_flat_geometry(H, W, C.byref(pitch), C.byref(blk)))
    blk = blk.value
    x = torch.randn(B * blk, 64, device=dev).to(dt)
    coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
    w = pack_conv3x3(torch.randn(64, 64, 3, 3, device=dev) / 24, dtype=dt)
    bias = torch.randn(64, device=dev)
    res = torch.randn(B * blk, 64, device=dev).to(dt) if res_mode == 1 else None
    out = torch.zeros(B * blk, 64, device=dev, dtype=torch.float32 if out_f32 else dt)
    st = torch.empty(B * blk // 128, 4, 16, 2, device=dev)

    def fn():
        L.check(lib.mcedm_conv_flat_fused(L.ptr(x), L.ptr(coef) if xf else None, L.ptr(w), L.ptr(bias), B, H, W, 64,
                                          L.ptr(out), out_f32, L.ptr(res), res_mode, 0, 0, 0, L.ptr(st), 1, L.stream_ptr()))
    px = B * H * W
    print(f"flat  {H}x{W} res={res_mode} xf={int(xf)} out_f32={out_f32}: " +
          timeit(fn, 2.0 * px * 64 * 576, px * 64 * 2.0 * (2 + (1 if res_mode else 0))), flush=True)


which = sys.argv[3] if len(sys.argv) > 3 else "all"
if which in ("all", "rows"):
    rows(128, 1, 0, 64, 0, False)
    rows(128, 1, 0, 64, 0, True)
    rows(128, 1, 0, 64, 1, True)
    rows(128, 2, 0, 32, 0, True)
    rows(128, 1, 2, 64, 0, True)
if which in ("all", "flat"):
    flat(64, 64, 0, False)
    flat(64, 64, 0, True)
    flat(64, 64, 1, True)
    flat(32, 32, 0, False)
    flat(32, 32, 1, True)
