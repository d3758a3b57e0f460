"""Parameter mirror of the reference DDPM U-Net `Model` (models/ddim_blocks.py:222-413) — SURVEY §8f rank 2, host half.

Same constructor argument (`hparams`), same sub-module names, shapes and registration order, and the same (torch
default) initialisation drawn in the same order, so `torch.manual_seed(s)` gives a bit-identical `state_dict` in both
code bases and reference checkpoints of the `config_ddim_res32*` / `config_edm_res32_cond_h` experiments load with
`strict=True` (tests/test_host_logic.py::test_ddpm_model_mirror_state_dict).

`forward` runs the sm_100a launch plan of `ddpm_engine.DdpmEngine` (the fused 16-bit convolution / attention kernels
of the ADM path + csrc/ddpm.cu: per-channel GroupNorm(32, eps 1e-6) statistics, temb folded into norm2's coefficients,
decimating Downsample, timestep MLP); the sub-modules below are parameter containers only.  Checked against the
fixtures from the unmodified reference (tests/golden/ddpm_path.pt) by tests/test_gpu_ddim.py.
"""
from __future__ import annotations

import torch
from torch import nn


def Normalize(in_channels):                                           # ddim_blocks.py:60-61
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):                                            # :64-80
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)


class Downsample(nn.Module):                                          # :83-103
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)


class ResnetBlock(nn.Module):                                         # :106-158
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)


class AttnBlock(nn.Module):                                           # :161-219
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)


class Model(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        m = hparams.model
        ch, out_channels = m.ch, m.out_ch
        channel_mult = tuple(m.ch_mult)
        attn_resolutions = m.attn_resolutions
        dropout = m.dropout
        cond_channels = m.cond_channels if hasattr(m, "cond_channels") else 0
        resolution = m.resolution
        resamp_with_conv = m.resamp_with_conv
        self.self_condition = m.self_cond if hasattr(m, "self_cond") else False
        self.cat_condition = m.cat_cond if hasattr(m, "cat_cond") else False
        self.dx_cond = m.dx_cond if hasattr(m, "dx_cond") else False
        self.cat_dx = m.cat_dx if hasattr(m, "cat_dx") else False
        if m.type == "bayesian":
            self.logvar = nn.Parameter(torch.zeros(hparams.diffusion.num_diffusion_timesteps))
        if self.dx_cond or (cond_channels > 0 and not self.cat_condition):
            raise NotImplementedError("separate condition / dx encoders of the DDPM U-Net are not mirrored")
        self.ch = ch
        self.channel_mult_emb = 4
        self.temb_ch = ch * self.channel_mult_emb
        self.num_resolutions = len(channel_mult)
        self.num_res_blocks = m.num_res_blocks
        self.resolution = resolution
        in_channels = m.in_channels * (2 if self.self_condition else 1)
        self.in_channels = in_channels + cond_channels if self.cat_condition else in_channels
        self.cond_channels = cond_channels

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        self.conv_in = nn.Conv2d(self.in_channels, ch, kernel_size=3, stride=1, padding=1)
        self.cond_enc = self.dx_enc = self.combine_enc = None

        curr_res = resolution
        in_ch_mult = (1,) + channel_mult
        self.down = nn.ModuleList()
        block_in = None
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * channel_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)

        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=dropout)

        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = skip_in = ch * channel_mult[i_level]
            for i_block in range(self.num_res_blocks + 1):
                if i_block == self.num_res_blocks:
                    skip_in = ch * in_ch_mult[i_level]
                block.append(ResnetBlock(in_channels=block_in + skip_in, out_channels=block_out,
                                         temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)

        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_channels, kernel_size=3, stride=1, padding=1)
        # what the samplers / engine read (same names as DhariwalUNet)
        self.attn_resolutions = list(attn_resolutions)
        self.resamp_with_conv = resamp_with_conv
        self.x_channels = m.in_channels
        self.out_channels = out_channels
        self.cat_channels = self.in_channels - m.in_channels          # self-conditioning copy + concatenated condition
        self._engine = None

    # the engine holds packed 16-bit weights and workspaces; it is not part of the module state
    def engine(self):
        if self._engine is None:
            from .ddpm_engine import DdpmEngine

            self._engine = DdpmEngine(self)
        return self._engine

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_engine" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engine"] = None
        return d

    def forward(self, x, t, cond=None, x_self_cond=None, dx=None):
        """Model.forward (ddim_blocks.py:415-470): x [B,C,128,128] fp32 CUDA, t [B] or [1] (timestep / c_noise),
        cond concatenated in front when `cat_cond`, x_self_cond (zeros when None) in front of x when `self_cond`
        (cat_conditioning, :378-390).  Returns [B,out_ch,128,128] fp32.  Inference only: no backward kernels."""
        if dx is not None:
            raise NotImplementedError("dx conditioning is not supported")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("the DDPM U-Net has an inference launch plan only (PlDdim's DDPM training loss is "
                                      "outside the hot path); call it in eval mode / under torch.no_grad()")
        cat_in = None
        if self.self_condition:
            cat_in = torch.zeros_like(x) if x_self_cond is None else x_self_cond
        elif x_self_cond is not None:
            raise ValueError("x_self_cond given to a network built without self_cond")
        if self.cat_condition and self.cond_channels > 0:
            if cond is None:
                cond = torch.zeros(x.shape[0], self.cond_channels, x.shape[2], x.shape[3], device=x.device, dtype=x.dtype)
            cat_in = cond if cat_in is None else torch.cat([cond, cat_in], dim=1)
        return self.engine().forward(x, t, cat_in)
