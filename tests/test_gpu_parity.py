"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI / the reference-shaped
Python surface, against the CPU oracle and the fixtures generated from the unmodified reference.

Tolerances (BASELINE.json north_star): per-step denoiser output within 1e-2 relative (bf16 tensor-core
operands, fp32 accumulation); integer / mask / indexing work bit-exact; fp64 sampler updates bit-exact
given the same network output.
"""
import copy

import pytest
import torch
import torch.nn.functional as F

from common import NoiseFeed, fixture_state, golden, hparams, stress_module, stress_unet
from mcedm_b200.utils import rel_l2
from oracle import edm_oracle as O

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def L():
    from mcedm_b200 import _lib

    _lib.lib()
    return _lib


# ----------------------------------------------------------------------------------------------- K1
def _pack(w, n_src):
    cout, cin, k, _ = w.shape
    segs, mats = [], []
    for i in range(n_src):
        for ky in range(k):
            for kx in range(k):
                segs.append((i, ky - k // 2, kx - k // 2))
                mats.append(w[:, 64 * i:64 * (i + 1), ky, kx])
    return torch.stack(mats, 0), segs


@pytest.mark.parametrize("B,H,W,n_src,k,N,res_mode,out_bf16", [
    (2, 128, 128, 1, 3, 64, 1, 0), (2, 128, 128, 2, 3, 64, 0, 0), (3, 64, 64, 1, 3, 64, 2, 0),
    (2, 64, 64, 1, 3, 64, 3, 0), (5, 32, 32, 2, 3, 64, 1, 0), (3, 32, 32, 1, 1, 192, 0, 1),
    (2, 64, 64, 1, 3, 128, 0, 0), (2, 128, 128, 1, 3, 16, 0, 0), (1, 16, 16, 1, 3, 64, 0, 0)])
def test_conv_igemm_matches_fp32_conv(L, dev, B, H, W, n_src, k, N, res_mode, out_bf16):
    lib = L.lib()
    g = torch.Generator().manual_seed(B * 1000 + H + N)
    srcs = [torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16).contiguous() for _ in range(n_src)]
    w = (torch.randn(N, 64 * n_src, k, k, generator=g) / (64 * n_src * k * k) ** 0.5).to(dev)
    wp, segs = _pack(w, n_src)
    wp = wp.to(torch.bfloat16).contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    res = {0: None, 1: (B, H, W, N), 2: (B, H // 2, W // 2, N), 3: (B, 2 * H, 2 * W, N)}[res_mode]
    res = torch.randn(*res, generator=g).to(dev) if res else None
    out = torch.full((B, H, W, N), float("nan"), device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    st = torch.full((B * H * W // 128, N // 4, 2), float("nan"), device=dev)
    L.check(lib.mcedm_conv_igemm(L.ptr_array(srcs), n_src, L.int_array([s[0] for s in segs]),
                                 L.int_array([s[1] for s in segs]), L.int_array([s[2] for s in segs]), len(segs),
                                 L.ptr(wp), L.ptr(bias), B, H, W, N, L.ptr(out), out_bf16, L.ptr(res), res_mode,
                                 L.ptr(st), 0, L.stream_ptr()), "conv_igemm")
    L.check_watchdog()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.cat([s.float() for s in srcs], -1).permute(0, 3, 1, 2)
    wq = torch.cat([wp[i * k * k:(i + 1) * k * k].float() for i in range(n_src)], 2)
    wq = wq.reshape(k, k, N, 64 * n_src).permute(2, 3, 0, 1).contiguous()
    ref = F.conv2d(x.double(), wq.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
    if res_mode == 1:
        ref = ref + res
    elif res_mode == 2:
        ref = ref + res.repeat_interleave(2, 1).repeat_interleave(2, 2)
    elif res_mode == 3:
        ref = ref + F.avg_pool2d(res.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel_l2(out.float(), ref) < (4e-3 if out_bf16 else 2e-6)
    if not out_bf16:
        v = out.reshape(-1, 128, N // 4, 4)
        assert torch.allclose(st[..., 0], v.sum(dim=(1, 3)), rtol=1e-4, atol=1e-3)
        assert torch.allclose(st[..., 1], (v * v).sum(dim=(1, 3)), rtol=1e-4, atol=1e-3)


def _to_flat(x_bhwc, pitch, blk):
    B, H, W, C = x_bhwc.shape
    flat = torch.zeros(B * blk, C, device=x_bhwc.device, dtype=x_bhwc.dtype)
    idx = (torch.arange(B, device=x_bhwc.device)[:, None, None] * blk
           + (torch.arange(H, device=x_bhwc.device)[None, :, None] + 1) * pitch
           + torch.arange(W, device=x_bhwc.device)[None, None, :])
    flat[idx.reshape(-1)] = x_bhwc.reshape(-1, C)
    return flat


@pytest.mark.parametrize("B,H,W,res_mode", [(3, 64, 64, 1), (2, 64, 64, 2), (2, 32, 32, 3), (5, 32, 32, 0),
                                             (40, 64, 64, 1), (2, 16, 16, 1)])
def test_conv_flat_and_conv_rows_match_conv_igemm_bitwise(L, dev, B, H, W, res_mode):
    """The row-resident (W=128) and padded-flat (W<=64) kernels issue the same MMAs in the same order as
    conv_igemm, so their results must be bit-identical to it (which is itself checked against fp64 above)."""
    import ctypes as C

    from mcedm_b200.engine import pack_conv3x3

    lib = L.lib()
    g = torch.Generator().manual_seed(H * 7 + res_mode)
    N = 64

    def run(Wc, Hc, flat):
        a = torch.randn(B, Hc, Wc, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
        w = pack_conv3x3((torch.randn(N, 64, 3, 3, generator=g) / 24).to(dev))
        bias = torch.randn(N, generator=g).to(dev)
        rshape = {0: None, 1: (B, Hc, Wc, N), 2: (B, Hc // 2, Wc // 2, N), 3: (B, 2 * Hc, 2 * Wc, N)}[res_mode]
        res = torch.randn(*rshape, generator=g).to(dev) if rshape else None
        segs = [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
        ref = torch.full((B, Hc, Wc, N), float("nan"), device=dev)
        st_ref = torch.empty(B * Hc * Wc // 128, 16, 2, device=dev)
        L.check(lib.mcedm_conv_igemm(L.ptr_array([a]), 1, L.int_array([s[0] for s in segs]),
                                     L.int_array([s[1] for s in segs]), L.int_array([s[2] for s in segs]), 9, L.ptr(w),
                                     L.ptr(bias), B, Hc, Wc, N, L.ptr(ref), 0, L.ptr(res), res_mode, L.ptr(st_ref),
                                     0, L.stream_ptr()))
        out = torch.full((B, Hc, Wc, N), float("nan"), device=dev)
        if flat:
            pitch, blk = C.c_int(0), C.c_int(0)
            L.check(lib.mcedm_flat_geometry(Hc, Wc, C.byref(pitch), C.byref(blk)))
            af = _to_flat(a, pitch.value, blk.value)
            st = torch.zeros(B * blk.value // 128, 4, 16, 2, device=dev)
            L.check(lib.mcedm_conv_flat(L.ptr(af), L.ptr(w), L.ptr(bias), B, Hc, Wc, N, L.ptr(out), L.ptr(res), res_mode,
                                        L.ptr(st), 0, L.stream_ptr()), "conv_flat")
        else:
            st = torch.zeros(B * Hc, 4, 16, 2, device=dev)
            L.check(lib.mcedm_conv_rows(L.ptr_array([a]), 1, None, 0, L.ptr(w), L.ptr(bias), B, Hc, N, L.ptr(out), 0,
                                        L.ptr(res), res_mode, L.ptr(st), 0, L.stream_ptr()), "conv_rows")
        L.check_watchdog()
        assert torch.equal(out, ref)
        tot, tot_ref = st.reshape(B, -1, 16, 2).sum(1), st_ref.reshape(B, -1, 16, 2).sum(1)
        assert torch.allclose(tot, tot_ref, rtol=1e-5, atol=1e-2)

    run(W, H, flat=True)
    if res_mode != 3:
        run(128, H, flat=False)


def test_umma_row_shifted_descriptor(L, dev):
    """Pins the hardware behaviour conv_rows / conv_flat rely on: a SWIZZLE_128B K-major operand addressed
    through a descriptor whose start is advanced by whole 128-byte rows (base_offset field 0) reads the
    row-shifted tile; and an MN-major B operand (attention's V) is consumed as stored."""
    lib = L.lib()
    g = torch.Generator().manual_seed(0)
    a = torch.randn(144, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
    bm = torch.randn(64, 64, generator=g).to(dev).to(torch.bfloat16).contiguous()
    for b_mn in (0, 1):
        for shift in (0, 1, 2, 7, 9, 16):
            out = torch.full((128, 64), float("nan"), device=dev)
            L.check(L.check_lib().mcedm_probe_umma(L.ptr(a), 144, L.ptr(bm), shift, 0, b_mn, L.ptr(out), L.stream_ptr()))
            L.check_watchdog()
            ref = a[shift:shift + 128].float() @ (bm.float() if b_mn else bm.float().t())
            assert (out - ref).abs().max().item() < 1e-4


# ----------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("B,H,W,rs,act,use_ss", [(2, 128, 128, 0, 1, True), (3, 32, 32, 1, 1, False),
                                                 (2, 64, 64, 2, 1, False), (2, 32, 32, 0, 0, False)])
def test_gn_apply_matches_group_norm(L, dev, B, H, W, rs, act, use_ss):
    lib = L.lib()
    g = torch.Generator().manual_seed(H + rs)
    x = (torch.randn(B, H, W, 64, generator=g) * 2 + 0.5).to(dev)
    gamma, beta = torch.randn(64, generator=g).to(dev), torch.randn(64, generator=g).to(dev)
    ss = (torch.randn(B, 128, generator=g) * 0.3).to(dev)
    st = torch.empty(B * H * W // 128, 16, 2, device=dev)
    L.check(lib.mcedm_gn_stats(L.ptr(x), B * H * W, L.ptr(st), L.stream_ptr()))
    Ho, Wo = (2 * H, 2 * W) if rs == 1 else (H // 2, W // 2) if rs == 2 else (H, W)
    out = torch.empty(B, Ho, Wo, 64, device=dev, dtype=torch.bfloat16)
    L.check(lib.mcedm_gn_apply(L.ptr(x), L.ptr(st), L.ptr(gamma), L.ptr(beta), L.ptr(ss) if use_ss else None, 128, 64,
                               1e-5, act, rs, B, H, W, 0, 0, 0, L.ptr(out), None, None, L.ptr(torch.empty(B, 128, device=dev)), 0, L.stream_ptr()))
    y = F.group_norm(x.permute(0, 3, 1, 2), 16, gamma, beta, 1e-5)
    if use_ss:
        y = torch.addcmul(ss[:, 64:, None, None], y, ss[:, :64, None, None] + 1)
    y = F.silu(y) if act else y
    y = y.repeat_interleave(2, 2).repeat_interleave(2, 3) if rs == 1 else F.avg_pool2d(y, 2) if rs == 2 else y
    assert rel_l2(out.float(), y.permute(0, 2, 3, 1)) < 4e-3      # one bf16 rounding of the output


# ----------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("B,Lq,scale", [(2, 1024, 1.0), (3, 256, 3.0), (1, 1024, 6.0)])
def test_attention_matches_fp32_softmax(L, dev, B, Lq, scale):
    lib = L.lib()
    g = torch.Generator().manual_seed(Lq)
    qkv = (torch.randn(B, Lq, 192, generator=g) * scale).to(dev).to(torch.bfloat16)
    out = torch.full((B, Lq, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out), None, 0, L.stream_ptr()), "attention")
    L.check_watchdog()
    q, k, v = qkv.double().split(64, dim=2)
    ref = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=2) @ v
    assert rel_l2(out.float(), ref) < 5e-3                         # bf16 P and bf16 output rounding
    chk = torch.empty(B, Lq, 64, device=dev)
    L.check(L.check_lib().mcedm_attention_ref(L.ptr(qkv), B, Lq, L.ptr(chk), L.stream_ptr()))
    assert rel_l2(chk, ref) < 1e-5


@pytest.mark.parametrize("fmt", [1, 0])
def test_attention_single_pass_overflow_falls_back_to_two_pass(L, dev, fmt):
    """Inference attention runs ONE pass over the keys with the first key block's row maximum as the softmax shift; a tile
    whose later keys score far above that (exponent argument > 15: beyond fp16's range) is flagged and redone by the
    two-pass kernel.  Scores here rise by ~+40 log2 units from key block 0 to the last one in sample 0 (every tile
    overflows), stay moderate in sample 1 (no tile does); both must match fp32 softmax, and so must the forced two-pass."""
    import os

    lib = L.lib()
    B, Lq = 2, 1024
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, Lq, 64, generator=g) * 1.5
    k = torch.randn(B, Lq, 64, generator=g)
    v = torch.randn(B, Lq, 64, generator=g)
    k[0] = k[0] * 0.05
    k[0, 896:] = q[0, :128] * 1.2              # the last key block aligns with the queries: scores ~ 1.2 |q|^2 ~ 170
    dt = torch.float16 if fmt else torch.bfloat16
    qkv = torch.cat([q, k, v], dim=2).to(dev).to(dt).contiguous()
    out = torch.full((B, Lq, 64), float("nan"), device=dev, dtype=dt)
    L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out), None, fmt, L.stream_ptr()), "attention")
    L.check_watchdog()
    qd, kd, vd = qkv.double().split(64, dim=2)
    ref = torch.softmax(qd @ kd.transpose(1, 2) / 8.0, dim=2) @ vd
    tol = 2e-3 if fmt else 8e-3
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out[0].float(), ref[0]) < tol and rel_l2(out[1].float(), ref[1]) < tol
    # with the log-sum-exp requested (training) the two-pass kernel runs alone: same answer
    lse = torch.empty(B, Lq, device=dev)
    out2 = torch.empty_like(out)
    L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out2), L.ptr(lse), fmt, L.stream_ptr()), "attention")
    assert rel_l2(out2.float(), ref) < tol
    assert rel_l2(out[1].float(), out2[1].float()) < tol


def test_attention_overflow_fallback_redoes_several_tiles_per_cta(L, dev):
    """The fallback launch is a 148-CTA grid striding over the tiles: at the sampler's batch sizes a CTA redoes MORE THAN
    ONE flagged tile (here tiles 0-7 of sample 0 and tiles 296-303 = 0-7 + 2 x 148 of sample 37 land on the same CTAs).
    Regression test: the kernel used to allocate TMEM per tile after relinquishing its allocation permit - a device
    exception on a CTA's second tile, first seen on one rank of an 8-GPU sampling run."""
    lib = L.lib()
    B, Lq = 40, 1024
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, Lq, 64, generator=g) * 1.5
    k = torch.randn(B, Lq, 64, generator=g)
    v = torch.randn(B, Lq, 64, generator=g)
    hot = [0, 20, 37]
    for s in hot:
        k[s] = k[s] * 0.05
        k[s, 896:] = q[s, :128] * 1.2
    qkv = torch.cat([q, k, v], dim=2).to(dev).half().contiguous()
    out = torch.full((B, Lq, 64), float("nan"), device=dev, dtype=torch.float16)
    for _ in range(2):
        L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(out), None, 1, L.stream_ptr()), "attention")
        torch.cuda.synchronize()
        L.check_watchdog()
    assert torch.isfinite(out.float()).all()
    for s in hot + [1, 39]:
        qd, kd, vd = qkv[s].double().split(64, dim=1)
        ref = torch.softmax(qd @ kd.t() / 8.0, dim=1) @ vd
        assert rel_l2(out[s].float(), ref) < 2e-3, s


# ----------------------------------------------------------------------------------------------- K5
def test_sampler_updates_bit_exact_against_torch_fp64(L, dev):
    lib = L.lib()
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 2, 128, 128
    noise = torch.randn(B, C, H, W, generator=g)
    cond = torch.randn(B, C, H, W, generator=g)
    mask = (torch.rand(B, C, H, W, generator=g) > 0.5).float()
    eps = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    F1, F2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    t_cur, t_next = torch.tensor(63.788, dtype=torch.float64), torch.tensor(56.7995, dtype=torch.float64)
    t_hat = t_cur + 0.3 * t_cur
    # torch (CPU, fp64/fp32) restatement of mcedm.py:594-628
    x0 = cond * (1 - mask) + (noise.double() * t_cur) * mask
    x_hat = x0 + (t_hat ** 2 - t_cur ** 2).sqrt() * 1 * eps * mask

    def D_of(xt, sig, Fx):
        cs, co, _, _ = O.precond_coeffs(sig)
        return cs * xt.float() + co * Fx

    d1 = D_of(x_hat, t_hat, F1)
    d_cur = (x_hat - d1.double()) / t_hat
    x_e = x_hat + (t_next - t_hat) * d_cur * mask
    d2 = D_of(x_e, t_next, F2)
    x_new = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * (x_e - d2.double()) / t_next) * mask
    # kernels
    from mcedm_b200.mcedm import precond_scalars

    # keep the device copies alive: L.ptr() only carries the address
    noise_d, cond_d, mask_d, eps_d, F1_d, F2_d = [t.to(dev).contiguous() for t in (noise, cond, mask, eps, F1, F2)]
    n = noise.numel()
    xg = torch.empty(B, C, H, W, device=dev, dtype=torch.float64)
    s = L.stream_ptr()
    L.check(lib.mcedm_edm_init(L.ptr(noise_d), L.ptr(cond_d), C, L.ptr(mask_d), float(t_cur), B, C, H, W, L.ptr(xg), s))
    assert torch.equal(xg.cpu(), x0)
    xh, xin = torch.empty_like(xg), torch.empty(B, C, H, W, device=dev)
    cs, co, ci, _ = precond_scalars(float(t_hat))
    cs2, co2, ci2, _ = precond_scalars(float(t_next))
    coef = float((t_hat ** 2 - t_cur ** 2).sqrt())
    L.check(lib.mcedm_edm_churn(L.ptr(xg), L.ptr(eps_d), L.ptr(mask_d), coef, ci, n, L.ptr(xh), L.ptr(xin), s))
    assert torch.equal(xh.cpu(), x_hat)
    assert torch.equal(xin.cpu(), O.precond_coeffs(t_hat)[2].reshape(()) * x_hat.float())
    dc, xe, Db = torch.empty_like(xg), torch.empty_like(xg), torch.empty_like(xin)
    L.check(lib.mcedm_edm_euler(L.ptr(xh), L.ptr(F1_d), L.ptr(mask_d), float(t_hat), float(t_next), cs, co, ci2, n,
                                L.ptr(dc), L.ptr(xe), L.ptr(xin), L.ptr(Db), s))
    assert torch.equal(Db.cpu(), d1) and torch.equal(dc.cpu(), d_cur) and torch.equal(xe.cpu(), x_e)
    xn = torch.empty_like(xg)
    L.check(lib.mcedm_edm_correct(L.ptr(xh), L.ptr(xe), L.ptr(F2_d), L.ptr(dc), L.ptr(mask_d), float(t_hat),
                                  float(t_next), cs2, co2, n, L.ptr(xn), L.ptr(Db), s))
    assert torch.equal(Db.cpu(), d2) and torch.equal(xn.cpu(), x_new)
    keep = mask == 0
    assert torch.equal(xn.cpu()[keep], cond.double()[keep])        # observed entries untouched, bit for bit


# ----------------------------------------------------------------------------------------------- fp16 operand format
def test_kernels_in_fp16_operand_format(L, dev):
    """op_fmt = 1: the same kernels with fp16 activations / weights / P (inference default): conv_rows, conv_flat,
    conv_igemm (fp16 output), gn_apply and attention against fp64 on the fp16-rounded operands."""
    import ctypes as C

    from mcedm_b200.engine import pack_conv3x3

    lib = L.lib()
    g = torch.Generator().manual_seed(21)
    B, N = 2, 64
    w4 = (torch.randn(N, 64, 3, 3, generator=g) / 24).to(dev).to(torch.float16)
    w = pack_conv3x3(w4.float(), dtype=torch.float16)
    for Hc, Wc in ((128, 128), (32, 32)):
        a = torch.randn(B, Hc, Wc, 64, generator=g).to(dev).to(torch.float16).contiguous()
        out = torch.full((B, Hc, Wc, N), float("nan"), device=dev)
        if Wc == 128:
            L.check(lib.mcedm_conv_rows(L.ptr_array([a]), 1, None, 0, L.ptr(w), None, B, Hc, N, L.ptr(out), 0, None, 0,
                                        None, 1, L.stream_ptr()), "conv_rows")
        else:
            pitch, blk = C.c_int(0), C.c_int(0)
            L.check(lib.mcedm_flat_geometry(Hc, Wc, C.byref(pitch), C.byref(blk)))
            L.check(lib.mcedm_conv_flat(L.ptr(_to_flat(a, pitch.value, blk.value)), L.ptr(w), None, B, Hc, Wc, N,
                                        L.ptr(out), None, 0, None, 1, L.stream_ptr()), "conv_flat")
        ref = F.conv2d(a.double().permute(0, 3, 1, 2), w4.double(), padding=1).permute(0, 2, 3, 1)
        assert rel_l2(out, ref) < 1e-5
    # 1x1 with fp16 output, then attention on it
    Lq = 256
    x = torch.randn(B, 16, 16, 64, generator=g).to(dev).to(torch.float16).contiguous()
    wq = (torch.randn(192, 64, generator=g) / 4).to(dev).to(torch.float16)
    qkv = torch.empty(B, Lq, 192, device=dev, dtype=torch.float16)
    L.check(lib.mcedm_conv_igemm(L.ptr_array([x]), 1, L.int_array([0]), L.int_array([0]), L.int_array([0]), 1,
                                 L.ptr(wq.reshape(1, 192, 64).contiguous()), None, B, 16, 16, 192, L.ptr(qkv), 1, None, 0,
                                 None, 1, L.stream_ptr()), "conv_igemm")
    ref_qkv = x.double().reshape(B, Lq, 64) @ wq.double().t()
    assert rel_l2(qkv.float(), ref_qkv) < 6e-4                     # one fp16 rounding of the output
    att = torch.empty(B, Lq, 64, device=dev, dtype=torch.float16)
    L.check(lib.mcedm_attention(L.ptr(qkv), B, Lq, L.ptr(att), None, 1, L.stream_ptr()), "attention")
    L.check_watchdog()
    q, k, v = qkv.double().split(64, dim=2)
    ref = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=2) @ v
    assert rel_l2(att.float(), ref) < 1e-3                          # fp16 P and fp16 output (bf16: 5e-3)
    # GroupNorm + SiLU -> fp16 operand
    xg = (torch.randn(B, 32, 32, 64, generator=g) * 2 + 0.5).to(dev)
    gamma, beta = torch.randn(64, generator=g).to(dev), torch.randn(64, generator=g).to(dev)
    st = torch.empty(B * 32 * 32 // 128, 16, 2, device=dev)
    L.check(lib.mcedm_gn_stats(L.ptr(xg), B * 32 * 32, L.ptr(st), L.stream_ptr()))
    o = torch.empty(B, 32, 32, 64, device=dev, dtype=torch.float16)
    L.check(lib.mcedm_gn_apply(L.ptr(xg), L.ptr(st), L.ptr(gamma), L.ptr(beta), None, 0, 64, 1e-5, 1, 0, B, 32, 32, 0, 0, 0,
                               L.ptr(o), None, None, L.ptr(torch.empty(B, 128, device=dev)), 1, L.stream_ptr()))
    y = F.silu(F.group_norm(xg.permute(0, 3, 1, 2), 16, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert rel_l2(o.float(), y) < 6e-4


def test_operand_format_error_budget(dev):
    """Network output error against the reference fixture in both operand formats: fp16 (inference default) sits an
    order of magnitude inside the 1e-2 bar; bf16 (what training uses) sits just inside it (operand rounding of ~30
    convolutions)."""
    g = golden("unet_forward.pt")
    net, cfg, _ = stress_unet()
    net = net.to(dev).eval()
    case = g["cases"][0]
    errs = {}
    with torch.no_grad():
        for fmt in (1, 0):
            net.engine().infer_fmt = fmt
            y = net(case["x"].to(dev), case["noise_labels"].to(dev), case["cond"].to(dev))
            errs[fmt] = rel_l2(y, case["out"])
    print("rel-L2 vs reference: fp16 operands %.3e, bf16 operands %.3e" % (errs[1], errs[0]))
    assert errs[1] < 2e-3 and errs[0] < BF16_TOL


# ----------------------------------------------------------------------------------------------- network
def test_unet_forward_within_bf16_bar_of_reference(dev):
    g = golden("unet_forward.pt")
    net, cfg, _ = stress_unet()
    net = net.to(dev)
    for case in g["cases"]:
        y = net(case["x"].to(dev), case["noise_labels"].to(dev), case["cond"].to(dev))
        assert y.shape == case["out"].shape and y.dtype == torch.float32
        assert rel_l2(y, case["out"]) < BF16_TOL
    # batch independence: a sample's output does not depend on its neighbours (GroupNorm / attention are per sample)
    c0 = g["cases"][1]
    y2 = net(c0["x"].to(dev), c0["noise_labels"].to(dev), c0["cond"].to(dev))
    y1 = net(c0["x"][1:].to(dev), c0["noise_labels"][1:].to(dev), c0["cond"][1:].to(dev))
    assert torch.equal(y2[1:], y1)


def test_cond_edm_network_within_bf16_bar(dev):
    g = golden("cond_edm_forward.pt")
    net, _, _ = stress_unet("config_adm_edm_res32_cond_h")
    net = net.to(dev)
    y = net(g["x"].to(dev), g["noise_labels"].to(dev), g["cond"].to(dev))
    assert rel_l2(y, g["out"]) < BF16_TOL


def test_cond_edm_sampler_and_training_step(dev):
    """Config 5 (`PlCondEdm`, config_adm_edm_res32_cond_h on Darcy-shaped fields): the kernel path behind the reference's
    single-task surface against the fixture produced by the unmodified reference (tests/golden/make_golden_cond.py):
    same RNG call sequence, per-evaluation D_x within the bar on the SAME input (teacher-forced through the oracle),
    final sample close, and the training-step loss."""
    from mcedm_b200 import data as D
    from mcedm_b200.cond_edm import PlCondEdm
    from mcedm_b200.utils import randomize_zero_init

    g = golden("cond_edm_path.pt")
    cfg = hparams("config_adm_edm_res32_cond_h")
    torch.manual_seed(1)
    pl = PlCondEdm(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    sd = {k: v.detach().clone() for k, v in pl.model.state_dict().items()}
    mcfg = dict(cfg.model.hparams.model)
    pl = pl.to(dev)
    st = g["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    a, u = D._FIELDS["darcy"](2, 128, first_seed=g["field_seed"])
    a, u = torch.from_numpy(a).to(dev), torch.from_numpy(u).to(dev)
    # ---- sampling (eval / no_grad, as Lightning runs test_step)
    pl.eval()
    state = pl.data_transform(a[:1], u[:1])
    h_n, u_n = state[..., :1], state[..., 1:2]
    feed = NoiseFeed(g["sample"]["seed"])
    pl._noise_hook = feed.hook
    pl._trace = []
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps = g["sample"]["steps"]
    u_noise = feed.draw(u_n)
    xs = pl.sample_edm(pl.get_cond_in(h_n, u_n, None, None), u_noise, sp, return_last=True, guide_dx=False)
    assert [tuple(c) for c in feed.calls] == [tuple(c) for c in g["sample"]["calls"]]
    assert xs.shape == (1, 1, 128, 128, 1) and xs.dtype == torch.float64
    assert len(pl._trace) == len(g["sample"]["denoised"])
    cond_c = h_n.permute(0, 3, 1, 2).contiguous().cpu()
    for (i, which, sigma, d, xt), ref in zip(pl._trace, g["sample"]["denoised"]):
        assert abs(sigma - ref["sigma"]) <= 1e-6 * max(1.0, ref["sigma"])
        with torch.no_grad():
            d_or, _ = O.denoise(sd, mcfg, xt.cpu(), torch.tensor(sigma, dtype=torch.float64), cond_c)
        assert rel_l2(d, d_or) < BF16_TOL
    # 5 chained evaluations on adversarial weights: the trajectories stay close (not a per-step bar)
    assert rel_l2(xs, g["sample"]["xs"]) < 5e-2
    pl._noise_hook, pl._trace = None, None
    # ---- training step (fp16 "fused16" training plan; loss within 2e-3)
    pl.train()
    pl.cond_p = 1.0
    nf = NoiseFeed(g["train"]["noise_seed"])
    pl._noise_hook = nf.hook
    torch.manual_seed(g["train"]["cpu_seed"])
    grid = torch.zeros(2, 128, 128, 1, device=dev)
    loss = pl.training_step((a, grid, grid, u), 0)
    assert abs(float(loss) - float(g["train"]["loss"])) < 2e-3 * abs(float(g["train"]["loss"]))
    loss.backward()
    gflat = pl.model.engine().flat_grad()
    gnorm = float(torch.sqrt((gflat.double() ** 2).sum()))
    assert abs(gnorm - float(g["train"]["grad_norm"])) < 5e-3 * float(g["train"]["grad_norm"])


def test_get_denoised_within_bf16_bar(dev):
    g = golden("denoise.pt")
    pl, _ = stress_module()
    pl = pl.to(dev)
    # get_denoised is the sampler's denoiser: the reference only calls it from validation_step / test_step / sample_edm,
    # which Lightning runs in eval mode under no_grad -> the inference plan (fp16 operands and activations, fused GN)
    pl.eval()
    for case in g["cases"]:
        with torch.no_grad():
            d, f = pl.get_denoised(pl.ema_model, case["xt"].to(dev), torch.tensor(case["sigma"], dtype=torch.float64),
                                   cond=case["cond"].to(dev), w=0.0)
        assert rel_l2(d, case["D"]) < BF16_TOL and rel_l2(f, case["F"]) < BF16_TOL
        assert rel_l2(f, case["F"]) < 3e-3, "the fp16 inference plan is expected well inside the bar"
    # the same call in train mode with autograd on goes through the TRAINING forward: the "fused16" plan runs the
    # inference data flow in fp16 (train16_engine.py), so it sits in the same 1.4e-3 class, far inside the 1e-2 bar
    # (the bf16 fp32-stream plan of round 1 measured 1.0e-2 here, AT the bar)
    pl.train()
    assert pl.ema_model.ma_model.engine().train_plan == "fused16"
    for case in g["cases"]:
        d, f = pl.get_denoised(pl.ema_model, case["xt"].to(dev), torch.tensor(case["sigma"], dtype=torch.float64),
                               cond=case["cond"].to(dev), w=0.0)
        assert rel_l2(f.detach(), case["F"]) < 3e-3 and rel_l2(d.detach(), case["D"]) < 3e-3


def test_sample_edm_trajectory_parity_with_injected_noise(dev):
    g = golden("trajectory.pt")
    pl, cfg = stress_module()
    pl = pl.to(dev)
    state, _, _, _ = fixture_state()
    for tr in g["trajs"]:
        feed = NoiseFeed(tr["seed"])
        pl._noise_hook = feed.hook
        pl._trace = []
        mask = tr["mask"].to(dev)
        cond_in = pl.get_cond_in(state.to(dev), mask, None, None).permute(0, 3, 1, 2).contiguous()
        mask_c = mask.permute(0, 3, 1, 2).contiguous()
        hu = feed.draw(mask_c)
        sp = copy.deepcopy(cfg.diff_sampler)
        sp.timesteps = tr["steps"]
        xs = pl.sample_edm(hu, cond_in, mask_c, sp, return_last=True, guide_dx=False)
        assert feed.calls == tr["calls"], "RNG draws differ from the reference in shape / dtype / order"
        assert len(pl._trace) == len(tr["denoised"])
        sd = {k: v.detach().cpu() for k, v in pl.ema_model.ma_model.state_dict().items()}
        mcfg = dict(cfg.model.hparams.model)
        for (i, which, sigma, d, xt), ref in zip(pl._trace, tr["denoised"]):
            assert abs(sigma - ref["sigma"]) <= 1e-6 * max(1.0, ref["sigma"])
            # per-step denoiser parity on the SAME input: the oracle (pinned to the reference) evaluated on the
            # state this path actually fed to the network
            with torch.no_grad():
                d_or, _ = O.denoise(sd, mcfg, xt.cpu(), torch.tensor(sigma, dtype=torch.float64), cond_in.cpu())
            print(f"step {i}.{which} sigma {sigma:9.4f}: rel-L2 vs oracle {rel_l2(d, d_or):.3e}")
            assert rel_l2(d, d_or) < BF16_TOL, f"step {i}.{which} sigma {sigma}"
            # and against the reference's own trajectory (inputs differ by the accumulated bf16 deviations)
            assert rel_l2(d, ref["D"]) < 3e-2, f"step {i}.{which} sigma {sigma} (trajectory)"
        assert xs.dtype == torch.float64 and tuple(xs.shape) == tuple(tr["xs"].shape)
        known = tr["mask"] == 0
        assert torch.equal(xs[:, -1].cpu()[known], state.double()[known])   # bit-identical observed entries
        assert rel_l2(xs, tr["xs"]) < 5e-2
    pl._noise_hook = pl._trace = None


def test_full_trajectory_rmse_within_2pct_of_reference(dev):
    g = golden("trajectory_full.pt")
    pl, cfg = stress_module()
    pl = pl.to(dev)
    state, h, u, _ = fixture_state()
    from mcedm_b200 import data as D

    mask = D.sample_mask(h[0], u[0], False)[g["mask_name"]].unsqueeze(0)
    feed = NoiseFeed(g["seed"])
    pl._noise_hook = feed.hook
    cond_in = pl.get_cond_in(state.to(dev), mask.to(dev), None, None).permute(0, 3, 1, 2).contiguous()
    mask_c = mask.permute(0, 3, 1, 2).contiguous().to(dev)
    hu = feed.draw(mask_c)
    xs = pl.sample_edm(hu, cond_in, mask_c, copy.deepcopy(cfg.diff_sampler), return_last=True)
    pl._noise_hook = None
    assert len(feed.calls) == g["n_calls"]
    rmse = torch.sqrt((((xs[:, -1].cpu() - state) * mask) ** 2).sum() / mask.sum())
    assert abs(float(rmse) - float(g["rmse"])) <= 0.02 * float(g["rmse"])


def test_no_cpu_fallback(dev):
    net, _, _ = stress_unet()
    with pytest.raises(Exception):
        net(torch.zeros(1, 2, 128, 128), torch.tensor([0.1]), torch.zeros(1, 2, 128, 128))


def test_fp32_accuracy_plan_within_1e4(dev):
    """north_star: per-step denoiser output within 1e-4 relative in fp32.  The split-operand plan (precise_engine.py:
    (hi, lo) fp16 operand pairs, three accumulating tensor-core launches per GEMM, fp32 attention) against the fixtures of
    the unmodified fp32 reference: network output F and denoiser output D at sigma in {80, 1.5, 0.05}."""
    g = golden("denoise.pt")
    pl, _ = stress_module()
    pl = pl.to(dev).eval()
    pl.ema_model.ma_model.engine().precision = "fp32"
    for case in g["cases"]:
        with torch.no_grad():
            d, f = pl.get_denoised(pl.ema_model, case["xt"].to(dev), torch.tensor(case["sigma"], dtype=torch.float64),
                                   cond=case["cond"].to(dev), w=0.0)
        assert rel_l2(f, case["F"]) < 1e-4 and rel_l2(d, case["D"]) < 1e-4
    gf = golden("unet_forward.pt")
    net, _, _ = stress_unet()
    net = net.to(dev).eval()
    net.engine().precision = "fp32"
    for case in gf["cases"]:
        with torch.no_grad():
            y = net(case["x"].to(dev), case["noise_labels"].to(dev), case["cond"].to(dev))
        assert rel_l2(y, case["out"]) < 1e-4


def test_full_size_micro_batch_properties(dev):
    """BASELINE configs[1] at its full micro-batch (256 rows of 128x128 fields, the bench's chunk), through properties
    that do not need the CPU oracle at that size: (1) rows are independent — a sub-batch sampled alone with the same
    draws gives bit-identical fields (this is what makes the 8-GPU row sharding exact); (2) observed entries
    (mask == 0) equal the condition bit for bit after the whole trajectory; (3) everything is finite."""
    pl, cfg = stress_module()
    pl = pl.to(dev).eval()
    B, lo, hi = 256, 40, 48
    g = torch.Generator().manual_seed(123)
    state = torch.randn(B, 2, 128, 128, generator=g)
    mask = torch.zeros(B, 2, 128, 128)
    mask[: B // 2, 1] = 1.0                       # first half: u missing; second half: h missing (mcedm.py eval masks)
    mask[B // 2:, 0] = 1.0
    cond = (state * (1 - mask) + torch.randn(B, 2, 128, 128, generator=g) * mask).to(dev)
    mask = mask.to(dev)
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps = 2
    draws = []

    def record(kind, like):
        t = torch.randn(like.shape, dtype=like.dtype, generator=g).to(like.device)
        draws.append(t)
        return t

    pl._noise_hook = record
    xs = pl.sample_edm(torch.zeros(B, 2, 128, 128, device=dev), cond, mask, sp, return_last=True)
    assert xs.shape == (B, 1, 128, 128, 2) and torch.isfinite(xs).all()
    keep = (mask == 0).permute(0, 2, 3, 1)
    assert torch.equal(xs[:, 0][keep], cond.permute(0, 2, 3, 1).double()[keep])
    replay = iter(draws)
    pl._noise_hook = lambda kind, like: next(replay)[lo:hi].contiguous()
    xs_sub = pl.sample_edm(torch.zeros(hi - lo, 2, 128, 128, device=dev), cond[lo:hi].contiguous(),
                           mask[lo:hi].contiguous(), sp, return_last=True)
    assert torch.equal(xs_sub, xs[lo:hi])


def test_precomputed_embedding_rows_are_bit_identical(dev):
    """The sampler evaluates the noise-embedding MLP + per-block affines for all of a trajectory's noise levels in one
    launch (UNetEngine.embedding_table) and hands each evaluation its row; the fields must be bit-identical to evaluating
    the MLP inside every evaluation, with and without CUDA-graph replay."""
    import copy

    from common import NoiseFeed

    pl, cfg = stress_module()
    pl = pl.to(dev).eval()
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps = 3
    g = torch.Generator().manual_seed(3)
    cond = torch.randn(2, 2, 128, 128, generator=g).to(dev)
    mask = torch.zeros(2, 2, 128, 128)
    mask[0, 1] = 1.0
    mask[1, 0] = 1.0
    mask = mask.to(dev)
    hu = torch.zeros(2, 2, 128, 128, device=dev)
    eng = pl.ema_model.ma_model.engine()
    outs = []
    for rows, graph in ((True, True), (False, True), (True, False)):
        eng.supports_ss_rows = rows
        pl.use_cuda_graph = graph
        pl._noise_hook = NoiseFeed(17).hook
        outs.append(pl.sample_edm(hu, cond, mask, sp))
    del eng.supports_ss_rows
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
