"""CPU oracle of the DDPM U-Net — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module.

Functional restatement (plain torch fp32 ops on a state_dict) of the network the reference builds for every experiment
whose `hparams.name` does not start with "adm" (models/ddim.py:40-43) — `config_ddim_res32*`, `config_edm_res32_cond_h`,
and therefore BASELINE config 4 as shipped:

    Model.forward / cat_conditioning        models/ddim_blocks.py:415-470, :366-390
    get_timestep_embedding, nonlinearity    :12-36
    ResnetBlock.forward                     :135-158   (GroupNorm(32, eps 1e-6) -> swish -> conv3x3 -> + temb_proj(swish(temb))
                                                        -> GroupNorm -> swish -> dropout(0) -> conv3x3; 1x1 nin_shortcut)
    AttnBlock.forward                       :191-219   (q, k, v 1x1 convs, softmax(q k / sqrt(C)), proj_out, residual)
    Downsample / Upsample                   :64-103    (pad (0,1,0,1) + stride-2 conv3x3; nearest x2 + conv3x3)

Options covered: cat_cond, self_cond, ch_mult / num_res_blocks / attn_resolutions as configured; not covered (raise):
the separate condition / dx encoders (`cat_cond: False` with cond_channels > 0, `dx_cond`), `type: bayesian`.
The arithmetic lives in PyTorch (see oracle/edm_oracle.py header).  Pinned to `tests/golden/ddpm_path.pt` (outputs of the
unmodified reference: one network evaluation and a 3-step RePaint-conditioned PlDdim.sample_edm trajectory,
tests/golden/make_golden_ddpm.py) by tests/test_oracle_golden.py.  No CUDA path exists for this network yet: this
oracle is the first step (scope order (a)) of that row.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e).type_as(t)
    e = t.float()[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    return F.pad(e, (0, 1, 0, 0)) if dim % 2 == 1 else e


def _swish(x):
    return x * torch.sigmoid(x)


def _norm(sd, pfx, x):
    return F.group_norm(x, 32, sd[pfx + ".weight"], sd[pfx + ".bias"], eps=1e-6)


def _conv(sd, pfx, x, stride=1, padding=1):
    return F.conv2d(x, sd[pfx + ".weight"], sd[pfx + ".bias"], stride=stride, padding=padding)


def _resnet(sd, pfx, x, temb):
    h = _conv(sd, pfx + ".conv1", _swish(_norm(sd, pfx + ".norm1", x)))
    h = h + F.linear(_swish(temb), sd[pfx + ".temb_proj.weight"], sd[pfx + ".temb_proj.bias"])[:, :, None, None]
    h = _conv(sd, pfx + ".conv2", _swish(_norm(sd, pfx + ".norm2", h)))
    if pfx + ".nin_shortcut.weight" in sd:
        x = _conv(sd, pfx + ".nin_shortcut", x, padding=0)
    elif pfx + ".conv_shortcut.weight" in sd:
        x = _conv(sd, pfx + ".conv_shortcut", x)
    return x + h


def _attn(sd, pfx, x):
    h = _norm(sd, pfx + ".norm", x)
    q, k, v = (_conv(sd, f"{pfx}.{n}", h, padding=0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, hh * ww), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, pfx + ".proj_out", h, padding=0)


def ddpm_unet_forward(sd: Dict[str, Tensor], model_cfg, x: Tensor, t: Tensor, cond: Optional[Tensor] = None,
                      x_self_cond: Optional[Tensor] = None) -> Tensor:
    """Model.forward(x, t, cond, x_self_cond) (ddim_blocks.py:415-470)."""
    if model_cfg.get("dx_cond", False) or (model_cfg.get("cond_channels", 0) > 0 and not model_cfg.get("cat_cond", False)) \
            or model_cfg.get("type") == "bayesian":
        raise NotImplementedError("separate condition / dx encoders and the bayesian variant are not restated")
    ch = model_cfg["ch"]
    mults = list(model_cfg["ch_mult"])
    nrb = model_cfg["num_res_blocks"]
    attn_res = list(model_cfg["attn_resolutions"])
    temb = timestep_embedding(t, ch)
    temb = F.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = F.linear(_swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])
    if model_cfg.get("self_cond", False):
        x = torch.cat([torch.zeros_like(x) if x_self_cond is None else x_self_cond, x], dim=1)
    cc = model_cfg.get("cond_channels", 0) if model_cfg.get("cat_cond", False) else 0
    if cc > 0:
        if cond is None:
            cond = torch.zeros(x.shape[0], cc, x.shape[2], x.shape[3], dtype=x.dtype)
        x = torch.cat([cond, x], dim=1)
    res = model_cfg["resolution"]
    hs = [_conv(sd, "conv_in", x)]
    for lvl in range(len(mults)):
        for i in range(nrb):
            h = _resnet(sd, f"down.{lvl}.block.{i}", hs[-1], temb)
            if res in attn_res:
                h = _attn(sd, f"down.{lvl}.attn.{i}", h)
            hs.append(h)
        if lvl != len(mults) - 1:
            if model_cfg.get("resamp_with_conv", True):
                hs.append(_conv(sd, f"down.{lvl}.downsample.conv", F.pad(hs[-1], (0, 1, 0, 1)), stride=2, padding=0))
            else:
                hs.append(F.avg_pool2d(hs[-1], kernel_size=2, stride=2))
            res //= 2
    h = _resnet(sd, "mid.block_1", hs[-1], temb)
    h = _attn(sd, "mid.attn_1", h)
    h = _resnet(sd, "mid.block_2", h, temb)
    for lvl in reversed(range(len(mults))):
        for i in range(nrb + 1):
            h = _resnet(sd, f"up.{lvl}.block.{i}", torch.cat([h, hs.pop()], dim=1), temb)
            if res in attn_res:
                h = _attn(sd, f"up.{lvl}.attn.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            if model_cfg.get("resamp_with_conv", True):
                h = _conv(sd, f"up.{lvl}.upsample.conv", h)
            res *= 2
    return _conv(sd, "conv_out", _swish(_norm(sd, "norm_out", h)))


def ddpm_net(sd, model_cfg, x, noise_labels, x_self_cond=None):
    """Adapter with the signature oracle.edm_oracle.vp_denoise / ddim_sample_with_repeat expect (cond = None)."""
    return ddpm_unet_forward(sd, model_cfg, x, noise_labels, x_self_cond=x_self_cond)
