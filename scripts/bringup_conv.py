"""GPU bring-up report for the tcgen05 conv kernel and the UMMA descriptor probes.

Run on a B200:  python scripts/bringup_conv.py  (writes gpurun_out/bringup_conv.json)
Every case is checked against torch (fp32 math on the bf16-rounded operands) and against the
CUDA-core direct kernel; nothing stops at the first failure so one GPU call reports everything.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
report = {"probe": [], "conv": []}


def probe():
    lib = L.lib()
    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.randn(144, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
    bm = torch.randn(64, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
    for b_mn in (0, 1):
        for shift in (0, 1, 2, 3, 7, 8, 9, 16):
            for bo in sorted({0, shift & 7}):
                out = torch.full((128, 64), float("nan"), device=dev)
                L.check(lib.mcedm_probe_umma(L.ptr(a), 144, L.ptr(bm), shift, bo, b_mn, L.ptr(out), L.stream_ptr()),
                        "probe")
                try:
                    L.check_watchdog()
                    wd = ""
                except Exception as e:  # noqa: BLE001
                    wd = str(e)
                A = a[shift:shift + 128].float()
                Bf = bm.float()
                ref = A @ (Bf if b_mn else Bf.t())
                err = (out - ref).abs().max().item()
                report["probe"].append(dict(b_mn=b_mn, shift=shift, base_offset=bo, max_err=err, watchdog=wd))
                print(f"probe b_mn={b_mn} shift={shift} bo={bo}: max_err={err:.3e} {wd}", flush=True)


def pack_weights(w, n_src):
    """[Cout, 64*n_src, k, k] -> ([n_seg][Cout][64] bf16, segs (src, dy, dx))"""
    cout, cin, k, _ = w.shape
    segs, mats = [], []
    for i in range(n_src):
        for ky in range(k):
            for kx in range(k):
                segs.append((i, ky - k // 2, kx - k // 2))
                mats.append(w[:, 64 * i:64 * (i + 1), ky, kx])
    return torch.stack(mats, 0), segs


def conv_case(name, B, H, W, n_src, k, N, res_mode=0, out_bf16=0, extra_skip_src=0, stats=True, seed=0):
    lib = L.lib()
    g = torch.Generator(device="cpu").manual_seed(seed)
    srcs = [torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16).contiguous() for _ in range(n_src)]
    w = (torch.randn(N, 64 * n_src, k, k, generator=g) / (64 * n_src * k * k) ** 0.5).to(dev)
    wp, segs = pack_weights(w, n_src)
    all_srcs = list(srcs)
    wskip = None
    if extra_skip_src:
        raw = [torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
               for _ in range(extra_skip_src)]
        wskip = (torch.randn(N, 64 * extra_skip_src, 1, 1, generator=g) / (64 * extra_skip_src) ** 0.5).to(dev)
        wp2, segs2 = pack_weights(wskip, extra_skip_src)
        segs += [(s + n_src, dy, dx) for (s, dy, dx) in segs2]
        wp = torch.cat([wp, wp2], 0)
        all_srcs += raw
    wp = wp.to(torch.bfloat16).contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    res = None
    if res_mode == 1:
        res = torch.randn(B, H, W, N, generator=g).to(dev)
    elif res_mode == 2:
        res = torch.randn(B, H // 2, W // 2, N, generator=g).to(dev)
    elif res_mode == 3:
        res = torch.randn(B, 2 * H, 2 * W, N, generator=g).to(dev)
    out = torch.full((B, H, W, N), float("nan"), device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    ntiles = B * H * W // 128
    st = torch.full((ntiles, N // 4, 2), float("nan"), device=dev) if stats else None
    t0 = time.time()
    rc = lib.mcedm_conv_igemm(L.ptr_array(all_srcs), len(all_srcs), L.int_array([s[0] for s in segs]),
                              L.int_array([s[1] for s in segs]), L.int_array([s[2] for s in segs]), len(segs),
                              L.ptr(wp), L.ptr(bias), B, H, W, N, L.ptr(out), out_bf16, L.ptr(res), res_mode,
                              L.ptr(st), L.stream_ptr())
    entry = dict(name=name, B=B, H=H, W=W, n_seg=len(segs), N=N, res_mode=res_mode, out_bf16=out_bf16)
    try:
        L.check(rc, name)
        L.check_watchdog()
    except Exception as e:  # noqa: BLE001
        entry["error"] = str(e)
        report["conv"].append(entry)
        print(name, "ERROR", e, flush=True)
        return
    torch.cuda.synchronize()
    entry["wall_ms"] = (time.time() - t0) * 1e3
    # torch reference on bf16-rounded operands, fp32 math, TF32 off
    x = torch.cat([s.float() for s in srcs], -1).permute(0, 3, 1, 2)
    wq = torch.cat([wp[i * k * k:(i + 1) * k * k].float() for i in range(n_src)], 2)  # [k*k, N, 64*n_src]
    wq = wq.reshape(k, k, N, 64 * n_src).permute(2, 3, 0, 1).contiguous()
    ref = torch.nn.functional.conv2d(x, wq, bias, padding=k // 2)
    if extra_skip_src:
        xr = torch.cat([s.float() for s in all_srcs[n_src:]], -1).permute(0, 3, 1, 2)
        ws = wp[n_src * k * k:].float()  # [extra, N, 64]
        ws = ws.permute(1, 0, 2).reshape(N, 64 * extra_skip_src, 1, 1)
        ref = ref + torch.nn.functional.conv2d(xr, ws)
    ref = ref.permute(0, 2, 3, 1)
    if res_mode == 1:
        ref = ref + res
    elif res_mode == 2:
        ref = ref + res.repeat_interleave(2, 1).repeat_interleave(2, 2)
    elif res_mode == 3:
        ref = ref + torch.nn.functional.avg_pool2d(res.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    o = out.float()
    err = (o - ref).abs().max().item()
    rel = ((o - ref).norm() / ref.norm()).item()
    entry.update(max_err=err, rel_l2=rel, nan=int(torch.isnan(o).sum().item()))
    # direct CUDA-core kernel
    seg_dev = torch.tensor(segs, dtype=torch.int32, device=dev).contiguous()
    out2 = torch.empty(B, H, W, N, device=dev)
    L.check(lib.mcedm_conv_direct_ref(L.ptr_array(all_srcs), len(all_srcs), L.ptr(seg_dev), len(segs), L.ptr(wp),
                                      L.ptr(bias), B, H, W, N, L.ptr(out2), L.ptr(res), res_mode, L.stream_ptr()),
            "direct")
    entry["direct_vs_torch"] = (out2 - ref).abs().max().item()
    if stats:
        v = (o if not out_bf16 else ref).reshape(ntiles, 128, N // 4, 4)
        s1 = v.sum(dim=(1, 3))
        s2 = (v * v).sum(dim=(1, 3))
        entry["stats_err"] = max((st[..., 0] - s1).abs().max().item() / (s1.abs().max().item() + 1e-6),
                                 (st[..., 1] - s2).abs().max().item() / (s2.abs().max().item() + 1e-6))
    report["conv"].append(entry)
    print(json.dumps(entry), flush=True)


def timing():
    """Device time of the 64->64 3x3 conv at 128x128, B=32 (the dominant layer class)."""
    lib = L.lib()
    B, H, W, N = 32, 128, 128, 64
    for n_src in (1, 2):
        srcs = [torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16) for _ in range(n_src)]
        w = torch.randn(N, 64 * n_src, 3, 3, device=dev) / 24
        wp, segs = pack_weights(w, n_src)
        wp = wp.to(torch.bfloat16).contiguous()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(B, H, W, N, device=dev)
        st = torch.empty(B * H * W // 128, N // 4, 2, device=dev)
        args = (L.ptr_array(srcs), n_src, L.int_array([s[0] for s in segs]), L.int_array([s[1] for s in segs]),
                L.int_array([s[2] for s in segs]), len(segs), L.ptr(wp), L.ptr(bias), B, H, W, N, L.ptr(out), 0, None,
                0, L.ptr(st), L.stream_ptr())
        for _ in range(3):
            L.check(lib.mcedm_conv_igemm(*args))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        iters = 20
        for _ in range(iters):
            L.check(lib.mcedm_conv_igemm(*args))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * B * H * W * N * 64 * n_src * 9
        entry = dict(name=f"timing_{64 * n_src}to64_128", ms=ms, tflops=flops / ms / 1e9)
        report["conv"].append(entry)
        print(json.dumps(entry), flush=True)


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    print(torch.cuda.get_device_name(0), flush=True)
    try:
        probe()
        conv_case("c64_128", 2, 128, 128, 1, 3, 64)
        conv_case("c64_128_res", 2, 128, 128, 1, 3, 64, res_mode=1)
        conv_case("c128_128", 2, 128, 128, 2, 3, 64, res_mode=0)
        conv_case("c64_128_skip2", 1, 128, 128, 1, 3, 64, extra_skip_src=2)
        conv_case("c64_64", 3, 64, 64, 1, 3, 64, res_mode=1)
        conv_case("c64_64_up", 2, 64, 64, 1, 3, 64, res_mode=2)
        conv_case("c64_64_down", 2, 64, 64, 1, 3, 64, res_mode=3)
        conv_case("c64_32", 5, 32, 32, 1, 3, 64, res_mode=1)
        conv_case("c128_32", 2, 32, 32, 2, 3, 64)
        conv_case("c64_16", 2, 16, 16, 1, 3, 64)
        conv_case("qkv_32", 3, 32, 32, 1, 1, 192, out_bf16=1)
        conv_case("proj_32", 3, 32, 32, 1, 1, 64, res_mode=1)
        conv_case("n128_64", 2, 64, 64, 1, 3, 128)
        conv_case("n16_128", 2, 128, 128, 1, 3, 16)
        conv_case("big_c64_128", 40, 128, 128, 1, 3, 64, res_mode=1)
        timing()
    finally:
        with open("gpurun_out/bringup_conv.json", "w") as f:
            json.dump(report, f, indent=1)
