"""GPU bring-up of the small kernels and the full U-Net forward against torch / the CPU oracle."""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L  # noqa: E402
from mcedm_b200.adm_blocks import DhariwalUNet  # noqa: E402
from mcedm_b200.config import compose  # noqa: E402
from mcedm_b200.utils import randomize_zero_init, rel_l2  # noqa: E402
from oracle import edm_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
rep = {}


def log(k, v):
    rep[k] = v
    print(k, v, flush=True)


def test_gn():
    lib = L.lib()
    g = torch.Generator().manual_seed(0)
    for (B, H, W, rs, act, use_ss) in [(2, 128, 128, 0, 1, True), (3, 32, 32, 1, 1, False), (2, 64, 64, 2, 1, False),
                                       (2, 32, 32, 0, 0, False)]:
        x = (torch.randn(B, H, W, 64, generator=g) * 2 + 0.5).to(dev)
        gamma = torch.randn(64, generator=g).to(dev)
        beta = torch.randn(64, generator=g).to(dev)
        ss = torch.randn(B, 128, generator=g).to(dev) * 0.3
        st = torch.empty(B * H * W // 128, 16, 2, device=dev)
        L.check(lib.mcedm_gn_stats(L.ptr(x), B * H * W, L.ptr(st), L.stream_ptr()))
        Ho, Wo = (2 * H, 2 * W) if rs == 1 else (H // 2, W // 2) if rs == 2 else (H, W)
        out = torch.empty(B, Ho, Wo, 64, device=dev, dtype=torch.bfloat16)
        raw = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16) if rs == 0 else None
        L.check(lib.mcedm_gn_apply(L.ptr(x), L.ptr(st), L.ptr(gamma), L.ptr(beta), L.ptr(ss) if use_ss else None, 128,
                                   64, 1e-5, act, rs, B, H, W, 0, 0, 0, L.ptr(out), L.ptr(raw), None, L.ptr(torch.empty(B, 128, device=dev)), L.stream_ptr()))
        xn = x.permute(0, 3, 1, 2)
        y = F.group_norm(xn, 16, gamma, beta, 1e-5)
        if use_ss:
            y = torch.addcmul(ss[:, 64:, None, None], y, ss[:, :64, None, None] + 1)
        if act:
            y = F.silu(y)
        if rs == 1:
            y = y.repeat_interleave(2, 2).repeat_interleave(2, 3)
        elif rs == 2:
            y = F.avg_pool2d(y, 2)
        y = y.permute(0, 2, 3, 1)
        err = (out.float() - y).abs().max().item()
        log(f"gn_{H}_{rs}_{act}", dict(max_err=err, ref_max=y.abs().max().item(),
                                      raw_err=(raw.float() - x).abs().max().item() if raw is not None else None))


def test_attn():
    lib = L.lib()
    g = torch.Generator().manual_seed(1)
    for (B, Lq, scale) in [(2, 1024, 1.0), (3, 256, 3.0), (1, 1024, 6.0)]:
        qkv = (torch.randn(B, Lq, 192, generator=g) * scale).to(dev).to(torch.bfloat16)
        out = torch.full((B, Lq, 64), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out), None, L.stream_ptr()), "attn")
        try:
            L.check_watchdog()
            wd = ""
        except Exception as e:  # noqa: BLE001
            wd = str(e)
        ref2 = torch.empty(B, Lq, 64, device=dev)
        L.check(lib.mcedm_attention_ref(L.ptr(qkv), B, Lq, L.ptr(ref2), L.stream_ptr()))
        q, k, v = qkv.float().split(64, dim=2)
        w = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=2)
        ref = w @ v
        log(f"attn_{B}_{Lq}_{scale}", dict(rel=rel_l2(out.float(), ref), max_err=(out.float() - ref).abs().max().item(),
                                           ref_kernel_rel=rel_l2(ref2, ref), wd=wd,
                                           nan=int(torch.isnan(out.float()).sum().item())))


def test_conv_in_emb():
    lib = L.lib()
    g = torch.Generator().manual_seed(2)
    B, H, W = 2, 128, 128
    x = torch.randn(B, 2, H, W, generator=g).to(dev)
    c = torch.randn(B, 2, H, W, generator=g).to(dev)
    w = (torch.randn(64, 4, 3, 3, generator=g) / 6).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    out = torch.empty(B, H, W, 64, device=dev)
    st = torch.empty(B * H * W // 128, 16, 2, device=dev)
    L.check(lib.mcedm_conv_in(L.ptr(x), 2, L.ptr(c), 2, L.ptr(w), L.ptr(b), B, H, W, L.ptr(out), L.ptr(st),
                              L.stream_ptr()))
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(torch.cat([c, x], 1), w, b, padding=1).permute(0, 2, 3, 1)
    v = out.reshape(-1, 128, 16, 4)
    log("conv_in", dict(max_err=(out - ref).abs().max().item(),
                        stats_err=(st[..., 0] - v.sum(dim=(1, 3))).abs().max().item()))


def test_unet():
    cfg = compose("config_adm_edm_mcedm_res32")
    hp = cfg.model.hparams
    torch.manual_seed(1)
    net = DhariwalUNet(hp)
    randomize_zero_init(net, 2)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev)
    g = torch.Generator().manual_seed(3)
    for (B, labels) in [(1, [0.3]), (2, [-1.0, 0.8]), (3, [1.0954])]:
        x = torch.randn(B, 2, 128, 128, generator=g)
        c = torch.randn(B, 2, 128, 128, generator=g)
        nl = torch.tensor(labels)
        t0 = time.time()
        y = net(x.to(dev), nl.to(dev), c.to(dev))
        torch.cuda.synchronize()
        t1 = time.time()
        try:
            L.check_watchdog()
            wd = ""
        except Exception as e:  # noqa: BLE001
            wd = str(e)
        with torch.no_grad():
            yo = O.unet_forward(sd, dict(hp.model), x, nl, c)
        log(f"unet_B{B}", dict(rel=rel_l2(y, yo), max_err=(y.cpu() - yo).abs().max().item(),
                               ref_max=yo.abs().max().item(), wall_ms=(t1 - t0) * 1e3, wd=wd,
                               nan=int(torch.isnan(y).sum().item())))
    # timing, B = 32
    B = 32
    x = torch.randn(B, 2, 128, 128, device=dev)
    c = torch.randn(B, 2, 128, 128, device=dev)
    nl = torch.tensor([0.5], device=dev)
    for _ in range(3):
        net(x, nl, c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net(x, nl, c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    log("unet_timing_B32", dict(ms=ms, us_per_sample=ms * 1e3 / B, tflops=18.797e9 * B / ms / 1e9))


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    try:
        for fn in (test_gn, test_attn, test_conv_in_emb, test_unet):
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                import traceback

                traceback.print_exc()
                log(fn.__name__ + "_EXC", str(e))
    finally:
        with open("gpurun_out/bringup_unet.json", "w") as f:
            json.dump(rep, f, indent=1)
