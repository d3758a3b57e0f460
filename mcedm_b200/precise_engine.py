"""fp32-accuracy inference plan: per-step denoiser output within 1e-4 of the fp32 reference (BASELINE north_star).

Same network as the other plans (DhariwalUNet.forward, models/adm_blocks.py:364-404), kept on the tensor cores: every
16-bit operand is a (hi, lo) fp16 pair, `x = hi + lo` to 22 significand bits, and every convolution / 1x1 projection is
issued as three accumulating launches  hi*W_hi + lo*W_hi + hi*W_lo  (fp32 accumulation; the dropped lo*W_lo term is 2^-22
relative) — SURVEY section 7, hard part 3: plain TF32 operands measure 1.1e-3 (fail), a three-term 16-bit split passes.
It runs on the unfused fp32-stream plan's kernels:

  * `mcedm_gn_apply_split` writes every GroupNorm(+scale/shift)+SiLU operand as a (hi, lo) pair (and the raw copies the
    1x1 skip projections read);
  * each term is a regular single-source launch of `mcedm_conv_rows` / `mcedm_conv_flat` / `mcedm_conv_igemm`; the
    first term carries bias + the block's residual, the others accumulate in place (res_mode 1 with res == out), the last
    one emits the GroupNorm statistics of the finished tensor;
  * attention is evaluated in fp32 on the CUDA cores (`mcedm_attention_f32`), as the reference's AttentionOp does.

Selected with `unet.engine().precision = "fp32"`.  ~4x the launches and ~3x the tensor work of the unfused plan: this is
the validation-grade mode, not the throughput path.
"""
from __future__ import annotations

import torch

from . import _lib as L


def split_packed(w_fp32_packed: torch.Tensor):
    """packed fp32 weights -> (hi, lo) fp16 tensors with hi + lo == w to 22 bits."""
    hi = w_fp32_packed.to(torch.float16)
    lo = (w_fp32_packed - hi.float()).to(torch.float16)
    return hi.contiguous(), lo.contiguous()


class PreciseMixin:
    # ------------------------------------------------------------------ weights
    def _pack_precise(self):
        from .engine import pack_conv3x3

        key = ("precise",) + self._param_key()
        if getattr(self, "_precise_key", None) == key:
            return
        u = self.unet
        f32 = torch.float32
        for b in self.blocks_enc + self.blocks_dec:
            m = b.mod
            b.p_w0 = [split_packed(pack_conv3x3(m.conv0.weight[:, 64 * i:64 * (i + 1)], dtype=f32)) for i in range(b.n_src)]
            b.p_w1 = split_packed(pack_conv3x3(m.conv1.weight, dtype=f32))
            if b.skip_conv:
                b.p_ws = [split_packed(pack_conv3x3(m.skip.weight[:, 64 * i:64 * (i + 1)], dtype=f32)) for i in range(b.n_src)]
            if b.attn:
                perm = torch.arange(192, device=m.qkv.weight.device).reshape(64, 3).t().reshape(-1)
                b.p_wqkv = split_packed(m.qkv.weight.detach()[perm].reshape(1, 192, 64).float())
                b.p_wproj = split_packed(m.proj.weight.detach().reshape(1, 64, 64).float())
        self.p_wout = split_packed(pack_conv3x3(u.out_conv.weight, n_out_pad=16, dtype=f32))
        self._precise_key = key

    # ------------------------------------------------------------------ launches
    def _gn_split(self, x, stats, parts, gamma, beta, ss, ss_stride, act, resample, B, Hin, Win, out_hi, out_lo, raw_hi,
                  raw_lo, st, eps, flat=None):
        pitch, blk = flat if flat is not None else (0, 0)
        coef = self._gn_coef.get((B, x.device.index))
        if coef is None:
            coef = self._gn_coef[(B, x.device.index)] = torch.empty(B, 128, device=x.device, dtype=torch.float32)
        L.check(self.lib.mcedm_gn_apply_split(L.ptr(x), L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(ss), ss_stride, 64, eps,
                                              act, resample, B, Hin, Win, parts, pitch, blk, L.ptr(out_hi), L.ptr(out_lo),
                                              L.ptr(raw_hi), L.ptr(raw_lo), L.ptr(coef), st), "gn_apply_split")
        L.LAUNCHES[0] += 1

    def _acc(self, terms, bias, B, H, W, N, out, res, res_mode, stats, st, flat=None):
        """out = bias + res + sum of terms; term = (kind, source, packed weights), kind "3x3" (source in the level's operand
        layout) or "1x1" (dense source).  Returns the number of statistics records per image written by the last term."""
        parts = 0
        for i, (kind, src, w) in enumerate(terms):
            first, last = i == 0, i == len(terms) - 1
            r, rm = (res, res_mode) if first else (out, 1)
            b_ = bias if first else None
            s_ = stats if last else None
            if kind == "3x3":
                parts = self._conv3x3([src], [], w, b_, B, H, W, N, out, r, rm, s_, st, flat=flat)
            else:
                self._conv([src], [(0, 0, 0)], w, b_, B, H, W, N, out, 0, r, rm, s_, st)
                parts = H * W // 128
        return parts

    @staticmethod
    def _three(kind, hi, lo, w):
        """the three product terms of (hi + lo) * (w_hi + w_lo) that matter"""
        return [(kind, hi, w[0]), (kind, lo, w[0]), (kind, hi, w[1])]

    def _pbuf(self, ws, name, like):
        t = ws["pool"].get(name)
        if t is None:
            t = ws["pool"][name] = torch.zeros_like(like)
        return t

    def _run_block_precise(self, blk, inputs, B, H_in, W_in, ws, emb_stride, st, dev):
        if blk.up:
            H, W, rs, res_mode = H_in * 2, W_in * 2, 1, 2
        elif blk.down:
            H, W, rs, res_mode = H_in // 2, W_in // 2, 2, 3
        else:
            H, W, rs, res_mode = H_in, W_in, 0, 1
        eps = blk.mod.norm0.eps
        flat = None
        a_bufs, a1_buf = ws["a"], ws["a1"]
        if self.use_flat and W <= 64:
            a_bufs, a1_buf, pitch, fblk = self._flat_buffers(ws, B, H, W, dev)
            flat = (pitch, fblk)
        a_lo = [self._pbuf(ws, ("p_a_lo", H, W, i), a_bufs[i]) for i in range(len(inputs))]
        a1_lo = self._pbuf(ws, ("p_a1_lo", H, W), a1_buf)
        terms0, skip_terms = [], []
        for i, (x, x_st, x_parts) in enumerate(inputs):
            raw_hi = ws["raw"][i] if blk.skip_conv else None
            raw_lo = self._pbuf(ws, ("p_raw_lo", i), ws["raw"][i]) if blk.skip_conv else None
            self._gn_split(x, x_st, x_parts, blk.g0[64 * i:64 * (i + 1)], blk.be0[64 * i:64 * (i + 1)], None, 0, 1, rs, B,
                           H_in, W_in, a_bufs[i], a_lo[i], raw_hi, raw_lo, st, eps, flat=flat)
            terms0 += self._three("3x3", a_bufs[i], a_lo[i], blk.p_w0[i])
            if blk.skip_conv:
                skip_terms += self._three("1x1", raw_hi, raw_lo, blk.p_ws[i])
        h_parts = self._acc(terms0, blk.b0, B, H, W, 64, ws["h"], None, 0, ws["h_st"], st, flat=flat)
        ss = ws["ss"][blk.aff_index * self._ss_rows * 128:]
        self._gn_split(ws["h"], ws["h_st"], h_parts, blk.g1, blk.be1, ss, emb_stride, 1, 0, B, H, W, a1_buf, a1_lo, None,
                       None, st, eps, flat=flat)
        out, out_st = self._tensor(ws, blk.name, B, H, W, dev)
        terms1 = skip_terms + self._three("3x3", a1_buf, a1_lo, blk.p_w1)       # a 3x3 term last: it writes the statistics
        if blk.skip_conv:
            parts = self._acc(terms1, blk.b1, B, H, W, 64, out, None, 0, out_st, st, flat=flat)
        else:
            parts = self._acc(terms1, blk.b1, B, H, W, 64, out, inputs[0][0], res_mode, out_st, st, flat=flat)
        if blk.attn:
            n2_lo = self._pbuf(ws, ("p_n2_lo",), ws["a1"])
            self._gn_split(out, out_st, parts, blk.g2, blk.be2, None, 0, 0, 0, B, H, W, ws["a1"], n2_lo, None, None, st, eps)
            pix = B * H * W
            qkv32 = ws["pool"].get("p_qkv32")
            if qkv32 is None:
                qkv32 = ws["pool"]["p_qkv32"] = torch.empty(pix * 192, device=dev, dtype=torch.float32)
                ws["pool"]["p_att32"] = torch.empty(pix * 64, device=dev, dtype=torch.float32)
                ws["pool"]["p_att_hi"] = torch.empty(pix * 64, device=dev, dtype=torch.float16)
                ws["pool"]["p_att_lo"] = torch.empty(pix * 64, device=dev, dtype=torch.float16)
            att32, att_hi, att_lo = ws["pool"]["p_att32"], ws["pool"]["p_att_hi"], ws["pool"]["p_att_lo"]
            self._acc(self._three("1x1", ws["a1"], n2_lo, blk.p_wqkv), blk.bqkv, B, H, W, 192, qkv32, None, 0, None, st)
            L.check(self.lib.mcedm_attention_f32(L.ptr(qkv32), B, H * W, L.ptr(att32), st), "attention_f32")
            L.check(self.lib.mcedm_split16(L.ptr(att32), pix * 64, L.ptr(att_hi), L.ptr(att_lo), st), "split16")
            out2, out2_st = self._tensor(ws, blk.name + ".attn", B, H, W, dev)
            parts = self._acc(self._three("1x1", att_hi, att_lo, blk.p_wproj), blk.bproj, B, H, W, 64, out2, out, 1, out2_st, st)
            out, out_st = out2, out2_st
        return (out, out_st, parts), H, W

    def _launch_all_precise(self, x, nl, cond, out):
        u = self.unet
        B, _, H, W = x.shape
        dev = x.device
        ws = self._workspace(B, H, W, dev)
        st = L.stream_ptr()
        lib = self.lib
        self._pack_precise()
        Bemb = nl.numel()
        self._ss_rows = Bemb
        emb_stride = 128 if Bemb == B and B > 1 else 0
        L.check(lib.mcedm_emb_mlp(L.ptr(nl), L.ptr(self.freqs), L.ptr(self.w_m0), L.ptr(self.b_m0), L.ptr(self.w_m1),
                                  L.ptr(self.b_m1), L.ptr(self.aff_w), L.ptr(self.aff_b), self.n_aff, Bemb, None,
                                  L.ptr(ws["ss"]), st), "emb_mlp")
        t0, t0_st = self._tensor(ws, "conv_in", B, H, W, dev)
        L.check(lib.mcedm_conv_in(L.ptr(x), u.x_channels, L.ptr(cond), u.cat_channels, L.ptr(self.w_in),
                                  L.ptr(self.b_in), B, H, W, L.ptr(t0), L.ptr(t0_st), st), "conv_in")      # fp32 FMA
        cur, ch, cw = (t0, t0_st, H * W // 128), H, W
        skips = [cur]
        for blk in self.blocks_enc:
            cur, ch, cw = self._run_block_precise(blk, [cur], B, ch, cw, ws, emb_stride, st, dev)
            skips.append(cur)
        for blk in self.blocks_dec:
            inputs = [cur]
            if blk.n_src == 2:
                inputs.append(skips.pop())
            cur, ch, cw = self._run_block_precise(blk, inputs, B, ch, cw, ws, emb_stride, st, dev)
        o_lo = self._pbuf(ws, ("p_out_lo",), ws["a1"])
        self._gn_split(cur[0], cur[1], cur[2], self.g_out, self.be_out, None, 0, 1, 0, B, H, W, ws["a1"], o_lo, None, None, st,
                       u.out_norm.eps)
        self._acc(self._three("3x3", ws["a1"], o_lo, self.p_wout), self.b_out, B, H, W, 16, ws["o16"], None, 0, None, st)
        L.check(lib.mcedm_head_to_nchw(L.ptr(ws["o16"]), 16, u.out_channels, B, H, W, L.ptr(out), st), "head_to_nchw")
        return out
