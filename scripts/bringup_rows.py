"""GPU bring-up of conv_rows (row-resident W=128 conv) against conv_igemm and torch."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
from mcedm_b200.engine import pack_conv3x3

dev = torch.device("cuda:0")
lib = L.lib()
SEG9 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]

def ev_time(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def case(name, B, H, n_halo, n_ctr, N, res_mode, timing=False):
    g = torch.Generator().manual_seed(hash(name) % 1000)
    W = 128
    halo = [torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16) for _ in range(n_halo)]
    ctr = [torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16) for _ in range(n_ctr)]
    w3 = (torch.randn(N, 64 * n_halo, 3, 3, generator=g) / (576 * n_halo) ** 0.5).to(dev)
    wp = pack_conv3x3(w3)
    if n_ctr:
        w1 = (torch.randn(N, 64 * n_ctr, 1, 1, generator=g) / (64 * n_ctr) ** 0.5).to(dev)
        wp = torch.cat([wp, pack_conv3x3(w1)], 0).contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    res = None
    if res_mode == 1: res = torch.randn(B, H, W, N, generator=g).to(dev)
    if res_mode == 2: res = torch.randn(B, H // 2, W // 2, N, generator=g).to(dev)
    out1 = torch.full((B, H, W, N), float("nan"), device=dev); st1 = torch.full((B * H, N // 4, 2), float("nan"), device=dev)
    out2 = torch.full((B, H, W, N), float("nan"), device=dev); st2 = torch.full((B * H, 4, N // 4, 2), float("nan"), device=dev)
    srcs = halo + ctr
    segs = [(i, dy, dx) for i in range(n_halo) for (dy, dx) in SEG9] + [(n_halo + i, 0, 0) for i in range(n_ctr)]
    def v1():
        L.check(lib.mcedm_conv_igemm(L.ptr_array(srcs), len(srcs), L.int_array([s[0] for s in segs]), L.int_array([s[1] for s in segs]),
            L.int_array([s[2] for s in segs]), len(segs), L.ptr(wp), L.ptr(bias), B, H, W, N, L.ptr(out1), 0, L.ptr(res), res_mode, L.ptr(st1), L.stream_ptr()))
    def v2():
        L.check(lib.mcedm_conv_rows(L.ptr_array(halo), n_halo, L.ptr_array(ctr) if ctr else None, n_ctr, L.ptr(wp), L.ptr(bias), B, H, N,
            L.ptr(out2), 0, L.ptr(res), res_mode, L.ptr(st2), L.stream_ptr()))
    e = dict(name=name)
    try:
        v1(); v2(); L.check_watchdog()
        e.update(max_diff=(out1 - out2).abs().max().item(), ref_max=out1.abs().max().item(), nan=int(torch.isnan(out2).sum()),
                 stats_diff=((st1 - st2.sum(1)).abs().max() / st1.abs().max()).item())
        if timing:
            flops = 2.0 * B * H * W * N * 64 * len(segs)
            t1, t2 = ev_time(v1), ev_time(v2)
            e.update(v1_ms=t1, v2_ms=t2, v1_tflops=flops / t1 / 1e9, v2_tflops=flops / t2 / 1e9)
    except Exception as ex:
        e["error"] = str(ex)
    print(json.dumps(e), flush=True)

case("r64", 2, 128, 1, 0, 64, 0)
case("r64_res", 3, 128, 1, 0, 64, 1)
case("r64_up", 2, 128, 1, 0, 64, 2)
case("r64_ctr2", 2, 128, 1, 2, 64, 0)
case("r16", 2, 128, 1, 0, 16, 0)
case("r64_h32", 5, 32, 1, 0, 64, 1)
case("r64_one", 1, 128, 1, 0, 64, 1)
case("t64_B8", 8, 128, 1, 0, 64, 1, timing=True)
case("t64_B32", 32, 128, 1, 0, 64, 1, timing=True)
case("t64_B128", 128, 128, 1, 0, 64, 0, timing=True)
case("t64ctr_B32", 32, 128, 1, 2, 64, 0, timing=True)
case("t16_B32", 32, 128, 1, 0, 16, 0, timing=True)
