"""Single-launch weight packing for the training step (mixin of `engine.UNetEngine`).

After every optimizer step the tensor-core operand copies of the parameters are stale: 16-bit `[segment][N][64]`
forward weights, flipped + transposed data-gradient weights, the permuted qkv weights, the padded head, and a few fp32
vectors that are not plain views of a parameter (conv1 bias + skip bias, the stacked `affine` matrices).  The torch
packing code (`UNetEngine.pack`, `TrainMixin.pack_train`) issues ~170 small permute / flip / cast kernels for this;
inside the training step that is 6-8 % of the step at 32 samples per GPU.

Every packed element is ONE parameter element (or zero padding, or for one bias the sum of two), so the whole re-pack is
a gather.  `PackPlan` discovers the gather indices by running the unchanged torch packing code once on parameters
whose values are their own flat indices (exact in fp32 below 2^24), lays all packed tensors out in two buffers
(16-bit, fp32), re-binds the engine's attributes to views of them, and from then on `run()` is one
`mcedm_pack_gather` launch (graph-capturable; bit-identical to the torch packing: same round-to-nearest casts).
"""
from __future__ import annotations

import torch

from . import _lib as L

_ALIGN = 512          # elements: every packed tensor starts on a 1 KiB (16-bit) / 2 KiB (fp32) boundary

# attributes that are real copies (not views of a parameter); lists hold one tensor per 64-channel input slice
_BLOCK_16 = ("w0", "w1", "wqkv", "wproj", "wd0", "wd1", "wdskip", "wdqkv", "wdproj")
_BLOCK_32 = ("bqkv",)
_ENGINE_16 = ("w_in_tc", "w_out", "wd_out")
_ENGINE_32 = ("b_out", "aff_w", "aff_b")


def _entries(eng):
    """[(owner, attr, list-index or None, is16)] of every packed copy that currently exists on the engine."""
    out = []

    def add(owner, name, is16):
        v = getattr(owner, name, None)
        if v is None:
            return
        if isinstance(v, (list, tuple)):
            out.extend((owner, name, i, is16) for i in range(len(v)))
        else:
            out.append((owner, name, None, is16))

    for b in eng.blocks_enc + eng.blocks_dec:
        for n in _BLOCK_16:
            add(b, n, True)
        for n in _BLOCK_32:
            add(b, n, False)
        if b.skip_conv:
            out.append((b, "b1", None, False))
    for n in _ENGINE_16:
        add(eng, n, True)
    for n in _ENGINE_32:
        add(eng, n, False)
    return out


def _get(owner, name, i):
    v = getattr(owner, name)
    return v if i is None else v[i]


def _set(owner, name, i, t):
    if i is None:
        setattr(owner, name, t)
    else:
        v = list(getattr(owner, name))
        v[i] = t
        setattr(owner, name, v)


class PackPlan:
    def __init__(self, eng):
        from . import train_engine

        self.eng = eng
        params = list(eng.unet.parameters())
        if not params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in params):
            raise L.McedmError("PackPlan needs contiguous fp32 CUDA parameters")
        if torch.cuda.is_current_stream_capturing():
            raise L.McedmError("PackPlan must be built outside CUDA-graph capture")
        dev = params[0].device
        n_total = sum(p.numel() for p in params)
        if n_total >= (1 << 24) - 1:
            raise NotImplementedError("PackPlan: index probing is exact only below 2^24 parameters")
        self.key = tuple(p.data_ptr() for p in params)
        base = min(self.key)
        self.base_ptr = base
        # flat position (1-based; 0 = padding) -> element offset from `base`
        table = torch.cat([torch.full((1,), -1, dtype=torch.int64, device=dev)] +
                          [(p.data_ptr() - base) // 4 + torch.arange(p.numel(), dtype=torch.int64, device=dev)
                           for p in params])
        # ---- 1. index mode: the torch packing code on index-valued parameters, every cast replaced by fp32
        saved = [p.data for p in params]
        off = 1
        for p in params:
            p.data = torch.arange(off, off + p.numel(), dtype=torch.float32, device=dev).view(p.shape)
            off += p.numel()
        fmt0 = eng._fmt
        self.fmt = eng.train_fmt
        try:
            eng._pack_dtype_override = torch.float32
            train_engine.DGRAD_DTYPE[0] = torch.float32
            eng._fmt = self.fmt
            eng.pack(force=True)
            eng.pack_train(force=True)
            probe = []
            for owner, name, i, is16 in _entries(eng):
                if name == "b1":
                    m = owner.mod
                    ia = m.conv1.bias.detach().round().long().reshape(-1)
                    ib = m.skip.bias.detach().round().long().reshape(-1)
                else:
                    ia = _get(owner, name, i).detach().float().round().long().reshape(-1)
                    ib = None
                probe.append((ia, ib))
        finally:
            for p, d in zip(params, saved):
                p.data = d
            eng._pack_dtype_override = None
            train_engine.DGRAD_DTYPE[0] = None
        # ---- 2. real mode: shapes / dtypes of the packed tensors, and the buffer layout
        eng._fmt = self.fmt
        eng.pack(force=True)
        eng.pack_train(force=True)
        self.entries = _entries(eng)
        assert len(self.entries) == len(probe)
        self.slots = []
        n16 = n32 = 0
        idx16, idx32a, idx32b = [], [], []
        for (owner, name, i, is16), (ia, ib) in zip(self.entries, probe):
            t = _get(owner, name, i)
            assert t.numel() == ia.numel() and (t.element_size() == 2) == is16, (name, t.shape, t.dtype)
            pad = (-t.numel()) % _ALIGN
            fill = torch.zeros(pad, dtype=torch.int64, device=dev)
            if is16:
                self.slots.append((n16, t.shape, t.dtype))
                idx16 += [ia, fill]
                n16 += t.numel() + pad
            else:
                self.slots.append((n32, t.shape, t.dtype))
                idx32a += [ia, fill]
                idx32b += [ib if ib is not None else torch.zeros_like(ia), fill]
                n32 += t.numel() + pad
        self.n16, self.n32 = n16, n32
        self.idx_a = table[torch.cat(idx16 + idx32a)].contiguous()
        self.idx_b = table[torch.cat(idx32b)].contiguous()
        self.dtype16 = torch.float16 if self.fmt else torch.bfloat16
        self.buf16 = torch.zeros(n16, dtype=self.dtype16, device=dev)
        self.buf32 = torch.zeros(n32, dtype=torch.float32, device=dev)
        self._base_holder = min(params, key=lambda p: p.data_ptr())    # keeps `base` alive
        eng._fmt = fmt0

    def valid_for(self, eng) -> bool:
        return self.key == tuple(p.data_ptr() for p in eng.unet.parameters()) and self.fmt == eng.train_fmt

    def bind(self):
        """Points the engine's packed attributes at the plan's buffers (cheap; no device work)."""
        for (owner, name, i, is16), (off, shape, dtype) in zip(self.entries, self.slots):
            buf = self.buf16 if is16 else self.buf32
            n = 1
            for s in shape:
                n *= s
            _set(owner, name, i, buf[off:off + n].view(shape))

    def run(self):
        import ctypes as C

        L.check(L.lib().mcedm_pack_gather(C.c_void_p(self.base_ptr), L.ptr(self.idx_a), L.ptr(self.idx_b), self.n16,
                                          self.n32, self.fmt, L.ptr(self.buf16), L.ptr(self.buf32), L.stream_ptr()),
                "pack_gather")


class PackMixin:
    def pack_fused(self):
        """Training-time replacement of `pack(force=True); pack_train(force=True)`: one gather launch, in the training
        plan's operand format (`train_fmt`: fp16 for the fused 16-bit plan, bf16 for the fp32-stream plan)."""
        plan = getattr(self, "_pack_plan", None)
        if plan is None or not plan.valid_for(self):
            plan = self._pack_plan = PackPlan(self)
        self._fmt = self.train_fmt
        plan.run()
        plan.bind()
        self._packed_key = self._param_key()
        self._packed_train_key = self._packed_key
