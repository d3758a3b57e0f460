"""Launch plan of the DDPM U-Net `Model` (models/ddim_blocks.py:415-470) on the sm_100a kernels — SURVEY §8f rank 2.

Same data flow as the fused ADM inference plan (fused_engine.py): every activation lives in HBM once, as a raw 16-bit
tensor (dense NHWC at 128x128, padded-flat at 64x64 / 32x32), and `swish(Normalize(x))` is applied by the consuming
convolution's transform warps from per-(sample, channel) coefficients.  What this network adds (csrc/ddpm.cu):

  * Normalize = GroupNorm(32, eps 1e-6): per-CHANNEL statistics from one pass over the stored tensor
    (`mcedm_gn_stats16`, cached per tensor: a skip tensor feeds two norms) and `mcedm_gn_coef_groups` for groups of
    2 (64 channels) or 4 (the decoder's 128-channel concat) channels;
  * `h + temb_proj(swish(temb))` between conv1 and norm2 (ResnetBlock.forward :140-146) is folded into norm2's
    coefficients (statistics corrected analytically): the sum never reaches HBM and conv1 keeps a batch-wide bias;
  * Downsample (:97-101) = the stride-1 conv on the raw tensor + `mcedm_decimate16` (odd positions);
    Upsample (:74-79) = `mcedm_gn_apply16` as a plain nearest-x2 copy (a = 1, b = 0, no activation) + conv;
  * q / k / v 1x1 convs stacked into one N = 192 GEMM, softmax(q k / 8) v on the attention kernel (C = 64);
  * the 256-wide timestep MLP + every block's temb_proj in one launch (`mcedm_ddpm_temb`).

Interface = what PlDdim's sampler needs from an engine: `_check_inputs`, `forward`, `forward_static` (CUDA-graph replay).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib as L
from .engine import UNetEngine, pack_conv3x3
from .fused_engine import Act, FusedMixin

N_SPLIT = 16      # position ranges per image of the statistics pass (fixed: the summation order must not depend on B)


class _Res:
    """Packed weights of one ResnetBlock."""

    def __init__(self, name, mod, index):
        self.name, self.mod, self.index = name, mod, index
        self.n_src = mod.in_channels // 64
        self.nin = hasattr(mod, "nin_shortcut")
        if hasattr(mod, "conv_shortcut"):
            raise NotImplementedError("conv_shortcut=True (3x3 shortcut) has no launch plan")

    def pack(self, dt):
        m = self.mod
        self.w1 = pack_conv3x3(m.conv1.weight, dtype=dt)
        self.b1 = m.conv1.bias.detach().float().contiguous()
        w2 = pack_conv3x3(m.conv2.weight, dtype=dt)
        b2 = m.conv2.bias.detach().float()
        if self.nin:
            w2 = torch.cat([w2, pack_conv3x3(m.nin_shortcut.weight, dtype=dt)], dim=0).contiguous()
            b2 = b2 + m.nin_shortcut.bias.detach().float()
        self.w2, self.b2 = w2, b2.contiguous()
        self.g1, self.be1 = m.norm1.weight.detach().float().contiguous(), m.norm1.bias.detach().float().contiguous()
        self.g2, self.be2 = m.norm2.weight.detach().float().contiguous(), m.norm2.bias.detach().float().contiguous()


class _Attn:
    def __init__(self, name, mod):
        self.name, self.mod = name, mod

    def pack(self, dt):
        m = self.mod
        self.wqkv = torch.cat([m.q.weight, m.k.weight, m.v.weight], 0).detach().reshape(1, 192, 64).to(dt).contiguous()
        self.bqkv = torch.cat([m.q.bias, m.k.bias, m.v.bias], 0).detach().float().contiguous()
        self.wproj = m.proj_out.weight.detach().reshape(1, 64, 64).to(dt).contiguous()
        self.bproj = m.proj_out.bias.detach().float().contiguous()
        self.g, self.be = m.norm.weight.detach().float().contiguous(), m.norm.bias.detach().float().contiguous()


class DdpmEngine(FusedMixin):
    _flat_geom = UNetEngine._flat_geom
    _conv = UNetEngine._conv
    _param_key = UNetEngine._param_key
    forward_static = UNetEngine.forward_static
    _forward_static = UNetEngine._forward_static
    forward = UNetEngine.forward

    def __init__(self, net):
        self.unet = net
        self.lib = L.lib()
        if net.ch != 64 or any(b.mod.out_channels != 64 for b in self._walk_resnets(net)):
            raise NotImplementedError("the DDPM launch plan is laid out for 64-channel levels (ch=64, ch_mult all 1)")
        if not net.resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False has no launch plan")
        self.res: List[_Res] = self._walk_resnets(net)
        self.attn = {}
        for lvl, d in enumerate(net.down):
            for i, a in enumerate(d.attn):
                self.attn[f"down.{lvl}.attn.{i}"] = _Attn(f"down.{lvl}.attn.{i}", a)
        self.attn["mid.attn_1"] = _Attn("mid.attn_1", net.mid.attn_1)
        for lvl, u in enumerate(net.up):
            for i, a in enumerate(u.attn):
                self.attn[f"up.{lvl}.attn.{i}"] = _Attn(f"up.{lvl}.attn.{i}", a)
        self.res_by_name = {r.name: r for r in self.res}
        self._ws, self._graphs = {}, {}
        self._packed_key = None
        self.infer_fmt = self._fmt = 1
        self.fused, self.precision = True, "fp16"

    @staticmethod
    def _walk_resnets(net) -> List[_Res]:
        """Every ResnetBlock in forward order (the order of the rows of mcedm_ddpm_temb's output)."""
        out = []
        for lvl, d in enumerate(net.down):
            for i, b in enumerate(d.block):
                out.append(_Res(f"down.{lvl}.block.{i}", b, len(out)))
        out.append(_Res("mid.block_1", net.mid.block_1, len(out)))
        out.append(_Res("mid.block_2", net.mid.block_2, len(out)))
        for lvl in reversed(range(len(net.up))):
            for i, b in enumerate(net.up[lvl].block):
                out.append(_Res(f"up.{lvl}.block.{i}", b, len(out)))
        return out

    # ------------------------------------------------------------------ weights
    def pack(self, force: bool = False):
        key = self._param_key()
        if not force and key == self._packed_key:
            return
        net = self.unet
        dev = net.conv_out.weight.device
        if dev.type != "cuda":
            raise L.McedmError("mcedm_b200 DDPM Model parameters must live on a CUDA (sm_100) device; there is no CPU path")
        dt = torch.float16
        with torch.no_grad():
            for r in self.res:
                r.pack(dt)
            for a in self.attn.values():
                a.pack(dt)
            cin = net.conv_in
            c_tot = cin.weight.shape[1]
            if 3 * c_tot > 15 or cin.weight.shape[0] != 64:
                raise NotImplementedError("conv_in with more than 5 input channels has no launch plan")
            wk = cin.weight.detach().float().permute(2, 0, 3, 1).reshape(3, 64, 3 * c_tot)          # ky, co, (kx, c)
            self.w_in_tc = torch.cat([wk, wk.new_zeros(3, 64, 64 - 3 * c_tot)], 2).to(dt).contiguous()
            self.b_in = cin.bias.detach().float().contiguous()
            self.w_out = pack_conv3x3(net.conv_out.weight, n_out_pad=16, dtype=dt)
            self.b_out = torch.cat([net.conv_out.bias.detach().float(),
                                    torch.zeros(16 - net.conv_out.weight.shape[0], device=dev)]).contiguous()
            self.g_out = net.norm_out.weight.detach().float().contiguous()
            self.be_out = net.norm_out.bias.detach().float().contiguous()
            self.ds = {f"down.{lvl}": (pack_conv3x3(d.downsample.conv.weight, dtype=dt),
                                       d.downsample.conv.bias.detach().float().contiguous())
                       for lvl, d in enumerate(net.down) if hasattr(d, "downsample")}
            self.us = {f"up.{lvl}": (pack_conv3x3(u.upsample.conv.weight, dtype=dt),
                                     u.upsample.conv.bias.detach().float().contiguous())
                       for lvl, u in enumerate(net.up) if hasattr(u, "upsample")}
            d0, d1 = net.temb.dense
            self.w_t0, self.b_t0 = d0.weight.detach().float().contiguous(), d0.bias.detach().float().contiguous()
            self.w_t1, self.b_t1 = d1.weight.detach().float().contiguous(), d1.bias.detach().float().contiguous()
            self.w_tp = torch.stack([r.mod.temb_proj.weight.detach().float() for r in self.res]).contiguous()
            self.b_tp = torch.stack([r.mod.temb_proj.bias.detach().float() for r in self.res]).contiguous()
        self._packed_key = key

    # ------------------------------------------------------------------ inputs
    def _check_inputs(self, x, noise_labels, cond):
        net = self.unet
        if not x.is_cuda:
            raise L.McedmError("Model.forward needs CUDA tensors: the sm_100a kernels have no CPU fallback")
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != net.x_channels:
            raise ValueError(f"x must be fp32 [B,{net.x_channels},H,W], got {x.dtype} {tuple(x.shape)}")
        B, _, H, W = x.shape
        if W != 128 or H % 4 != 0:
            raise ValueError(f"unsupported field size {H}x{W} (the 16-bit launch plan is laid out for 128-wide fields)")
        x = x.contiguous()
        if net.cat_channels > 0:
            if cond is None:
                cond = torch.zeros(B, net.cat_channels, H, W, device=x.device, dtype=torch.float32)
            if cond.shape != (B, net.cat_channels, H, W) or cond.dtype != torch.float32:
                raise ValueError(f"concatenated input must be fp32 [B,{net.cat_channels},H,W], got {tuple(cond.shape)}")
            cond = cond.contiguous()
        else:
            cond = None
        nl = noise_labels.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        if nl.numel() not in (1, B):
            raise ValueError(f"t must have 1 or B={B} entries, got {nl.numel()}")
        return x, nl, cond

    # ------------------------------------------------------------------ launches
    def _coef(self, ws, cache, key, act: Act, gamma, beta, cpg, eps, B, st, shift=None, shift_stride=0):
        """(a | b) of swish(GroupNorm(act [+ shift])); the per-channel statistics of a tensor are computed once."""
        part = cache.get(id(act))
        if part is None:
            part = cache[id(act)] = self._fbuf(ws, "st16." + key, (B, N_SPLIT, 64, 2), torch.float32, act.t.device)
            npos = act.flat[1] if act.flat is not None else act.H * act.W
            L.check(self.lib.mcedm_gn_stats16(L.ptr(act.t), npos, B, self._fmt, N_SPLIT, L.ptr(part), st), "gn_stats16")
        coef = self._fbuf(ws, "coef." + key, (B, 128), torch.float32, act.t.device)
        L.check(self.lib.mcedm_gn_coef_groups(L.ptr(part), N_SPLIT, act.H * act.W, L.ptr(gamma), L.ptr(beta), cpg, eps,
                                              L.ptr(shift), shift_stride, B, L.ptr(coef), st), "gn_coef_groups")
        return coef

    def _resnet(self, r: _Res, inputs: List[Act], B, ws, cache, tb, tb_stride, st, dev) -> Act:
        H, W = inputs[0].H, inputs[0].W
        eps = r.mod.norm1.eps
        cpg = 2 * len(inputs)                                         # 32 groups over 64 / 128 channels
        coefs = [self._coef(ws, cache, f"{r.name}.n1.{i}", a, r.g1[64 * i:64 * (i + 1)], r.be1[64 * i:64 * (i + 1)], cpg, eps,
                            B, st) for i, a in enumerate(inputs)]
        h = self._fact(ws, f"h.{H}", B, H, W, dev, stats=False)
        self._fconv(inputs, coefs, r.w1, r.b1, B, h, None, 0, st, ws=ws, stats=False)
        cache.pop(id(h), None)                                       # h.{H} is rewritten by every block of the level
        coef2 = self._coef(ws, cache, f"{r.name}.n2", h, r.g2, r.be2, 2, r.mod.norm2.eps, B, st, shift=tb[r.index],
                           shift_stride=tb_stride)
        cache.pop(id(h), None)
        out = self._fact(ws, r.name, B, H, W, dev, stats=False)
        if r.nin:
            self._fconv([h], [coef2], r.w2, r.b2, B, out, None, 0, st, ctr=inputs, ws=ws, stats=False)
        else:
            self._fconv([h], [coef2], r.w2, r.b2, B, out, inputs[0], 1, st, ws=ws, stats=False)
        return out

    def _attn_block(self, a: _Attn, x: Act, B, ws, cache, st, dev) -> Act:
        H, W = x.H, x.W
        coef = self._coef(ws, cache, a.name + ".n", x, a.g, a.be, 2, a.mod.norm.eps, B, st)
        a2 = self._fbuf(ws, "att.in", (B, H, W, 64), self._dt16(), dev)
        self._fapply16(x, coef, 0, 0, B, None, st, dense_out=a2)
        qkv = self._fbuf(ws, "att.qkv", (B, H * W, 192), self._dt16(), dev)
        att = self._fbuf(ws, "att.out", (B, H * W, 64), self._dt16(), dev)
        self._conv([a2], [(0, 0, 0)], a.wqkv, a.bqkv, B, H, W, 192, qkv, 1, None, 0, None, st)
        L.check(self.lib.mcedm_attention(L.ptr(qkv), B, H * W, L.ptr(att), None, self._fmt, st), "attention")
        L.LAUNCHES[0] += 1                       # single-pass kernel + the flagged-tile two-pass launch
        out = self._fact(ws, a.name, B, H, W, dev, stats=False)
        pitch, fblk = out.flat if out.flat is not None else (0, 0)
        L.check(self.lib.mcedm_conv_igemm16(L.ptr_array([att]), 1, L.int_array([0]), L.int_array([0]), L.int_array([0]), 1,
                                            L.ptr(a.wproj), L.ptr(a.bproj), B, H, W, 64, L.ptr(out.t), L.ptr(x.t), 1, pitch,
                                            fblk, None, self._fmt, st), "conv_igemm16")
        return out

    def _launch_all(self, x, nl, cond, out):
        net = self.unet
        B, _, H, W = x.shape
        dev = x.device
        ws = self._fws(B, H, W, dev)
        st = L.stream_ptr()
        lib = self.lib
        cache = {}
        Bt = nl.numel()
        tb_all = self._fbuf(ws, "temb", (len(self.res), Bt, 64), torch.float32, dev)
        L.check(lib.mcedm_ddpm_temb(L.ptr(nl), Bt, L.ptr(self.w_t0), L.ptr(self.b_t0), L.ptr(self.w_t1), L.ptr(self.b_t1),
                                    L.ptr(self.w_tp), L.ptr(self.b_tp), len(self.res), L.ptr(tb_all), st), "ddpm_temb")
        tb_stride = 64 if (Bt == B and B > 1) else 0
        ident = ws.get("ident")
        if ident is None:
            ident = ws["ident"] = torch.cat([torch.ones(B, 64, device=dev), torch.zeros(B, 64, device=dev)], 1).contiguous()
        t0 = self._fact(ws, "conv_in", B, H, W, dev)
        L.check(lib.mcedm_conv_in_tc16(L.ptr(x), net.x_channels, L.ptr(cond), net.cat_channels, L.ptr(self.w_in_tc),
                                       L.ptr(self.b_in), B, H, L.ptr(t0.t), L.ptr(t0.st), self._fmt, st), "conv_in_tc16")
        hs = [t0]
        res = net.resolution
        n_lvl = len(net.down)
        for lvl in range(n_lvl):
            for i in range(net.num_res_blocks):
                h = self._resnet(self.res_by_name[f"down.{lvl}.block.{i}"], [hs[-1]], B, ws, cache, tb_all, tb_stride, st, dev)
                if res in net.attn_resolutions:
                    h = self._attn_block(self.attn[f"down.{lvl}.attn.{i}"], h, B, ws, cache, st, dev)
                hs.append(h)
            if lvl != n_lvl - 1:
                src = hs[-1]
                w, b = self.ds[f"down.{lvl}"]
                tmp = self._fact(ws, f"ds.tmp.{src.H}", B, src.H, src.W, dev, stats=False)
                self._fconv([src], None, w, b, B, tmp, None, 0, st, ws=ws, stats=False)       # raw input: no transform
                dn = self._fact(ws, f"down.{lvl}.ds", B, src.H // 2, src.W // 2, dev, stats=False)
                ip, ib = tmp.flat if tmp.flat is not None else (0, 0)
                op, ob = dn.flat if dn.flat is not None else (0, 0)
                L.check(lib.mcedm_decimate16(L.ptr(tmp.t), ip, ib, B, src.H, src.W, L.ptr(dn.t), op, ob, st), "decimate16")
                hs.append(dn)
                res //= 2
        h = self._resnet(self.res_by_name["mid.block_1"], [hs[-1]], B, ws, cache, tb_all, tb_stride, st, dev)
        h = self._attn_block(self.attn["mid.attn_1"], h, B, ws, cache, st, dev)
        h = self._resnet(self.res_by_name["mid.block_2"], [h], B, ws, cache, tb_all, tb_stride, st, dev)
        for lvl in reversed(range(n_lvl)):
            for i in range(net.num_res_blocks + 1):
                h = self._resnet(self.res_by_name[f"up.{lvl}.block.{i}"], [h, hs.pop()], B, ws, cache, tb_all, tb_stride, st,
                                 dev)
                if res in net.attn_resolutions:
                    h = self._attn_block(self.attn[f"up.{lvl}.attn.{i}"], h, B, ws, cache, st, dev)
            if lvl != 0:
                w, b = self.us[f"up.{lvl}"]
                opnd = self._fact(ws, f"op.{h.H * 2}", B, h.H * 2, h.W * 2, dev, stats=False)
                self._fapply16(h, ident, 0, 1, B, opnd, st)          # nearest x2 of the raw tensor (a = 1, b = 0, no act)
                up = self._fact(ws, f"up.{lvl}.us", B, h.H * 2, h.W * 2, dev, stats=False)
                self._fconv([opnd], None, w, b, B, up, None, 0, st, ws=ws, stats=False)
                h = up
                res *= 2
        coef = self._coef(ws, cache, "out", h, self.g_out, self.be_out, 2, net.norm_out.eps, B, st)
        L.check(lib.mcedm_conv_head_fused(L.ptr(h.t), L.ptr(coef), L.ptr(self.w_out), L.ptr(self.b_out), B, H,
                                          net.out_channels, L.ptr(out), self._fmt, st), "conv_head_fused")
        return out
