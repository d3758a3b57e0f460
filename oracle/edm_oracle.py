"""CPU oracle of the m-cedm hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; nothing under `mcedm_b200/` does.

What it is: a functional restatement (plain `torch` fp32/fp64 ops on a state_dict, no nn.Modules, no
Lightning) of the reference algorithm for the path named by BASELINE.json.north_star:

    U-Net forward          models/adm_blocks.py:364-404 (DhariwalUNet.forward), :159-181 (UNetBlock),
                           :36-82 (Conv2d incl. up/down resampling), :86-97 (GroupNorm),
                           :103-109 (AttentionOp), :185-199 (PositionalEmbedding), :19-32 (Linear)
    EDM preconditioning    models/mcedm.py:199-211 (model_precond), :443-461 (get_denoised)
    training loss          models/mcedm.py:213-239, :254-281 ; models/losses.py:39-59
    Heun sampler + masks   models/mcedm.py:570-638 (sample_edm), :241-252 (get_cond_in)
    eval metric            models/losses.py:62-78 (MaskedLoss)
    mask generators        datamodules/h5_dataset.py:232-255, :306-393

The arithmetic itself lives in a third-party dependency that is not vendored under the reference:
PyTorch (pinned `pytorch=1.13.1` in requirements.yml:8; this image has 2.11.0).  The oracle therefore
calls the same torch primitives the reference calls (F.conv2d, F.group_norm, softmax, ...).

Pinning: the reference ships NO tests and NO golden vectors for this path ("parity unpinned" by the
reference itself).  The oracle is instead pinned against outputs of the reference
itself, produced in the build container by importing /root/reference unmodified
(tests/golden/make_golden.py -> tests/golden/*.pt); tests/test_oracle_golden.py checks every function
here against those fixtures.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# structure of the network (derived from hparams exactly as DhariwalUNet.__init__, adm_blocks.py:277-317)
# ------------------------------------------------------------------------------------------------
def unet_plan(model_cfg) -> Dict[str, List[dict]]:
    ch = model_cfg["ch"]
    mults = list(model_cfg["ch_mult"])
    res0 = model_cfg["resolution"]
    nrb = model_cfg["num_res_blocks"]
    attn = list(model_cfg["attn_resolutions"])
    enc, dec = [], []
    cout = None
    for level, mult in enumerate(mults):
        res = res0 >> level
        if level == 0:
            cout = ch * mult
            enc.append(dict(name=f"{res}x{res}_conv", kind="conv", cout=cout))
        else:
            enc.append(dict(name=f"{res}x{res}_down", kind="block", cin=cout, cout=cout, down=True, up=False,
                            attn=False))
        for idx in range(nrb):
            cin, cout = cout, ch * mult
            enc.append(dict(name=f"{res}x{res}_block{idx}", kind="block", cin=cin, cout=cout, down=False, up=False,
                            attn=res in attn))
    skips = [b["cout"] for b in enc]
    for level, mult in reversed(list(enumerate(mults))):
        res = res0 >> level
        if level == len(mults) - 1:
            dec.append(dict(name=f"{res}x{res}_in0", kind="block", cin=cout, cout=cout, down=False, up=False,
                            attn=True))
            dec.append(dict(name=f"{res}x{res}_in1", kind="block", cin=cout, cout=cout, down=False, up=False,
                            attn=False))
        else:
            dec.append(dict(name=f"{res}x{res}_up", kind="block", cin=cout, cout=cout, down=False, up=True,
                            attn=False))
        for idx in range(nrb + 1):
            cin = cout + skips.pop()
            cout = ch * mult
            dec.append(dict(name=f"{res}x{res}_block{idx}", kind="block", cin=cin, cout=cout, down=False, up=False,
                            attn=res in attn))
    return dict(enc=enc, dec=dec)


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
def _conv(x: Tensor, w: Optional[Tensor], b: Optional[Tensor], up=False, down=False) -> Tensor:
    """Conv2d.forward with resample_filter=[1,1] (adm_blocks.py:57-82): the resampling filter is the
    2x2 box [[.25,.25],[.25,.25]]; `up` applies it x4 as a stride-2 transposed conv (= nearest x2),
    `down` as a stride-2 depthwise conv (= 2x2 mean); both happen BEFORE the kxk filter."""
    c = x.shape[1]
    if up:
        f = torch.full((c, 1, 2, 2), 1.0, dtype=x.dtype, device=x.device)
        x = F.conv_transpose2d(x, f, groups=c, stride=2)
    if down:
        f = torch.full((c, 1, 2, 2), 0.25, dtype=x.dtype, device=x.device)
        x = F.conv2d(x, f, groups=c, stride=2)
    if w is not None:
        x = F.conv2d(x, w.to(x.dtype), padding=w.shape[-1] // 2)
    if b is not None:
        x = x + b.to(x.dtype).reshape(1, -1, 1, 1)
    return x


def _gn(x: Tensor, w: Tensor, b: Tensor, eps=1e-5) -> Tensor:
    groups = min(32, x.shape[1] // 4)  # adm_blocks.py:87-90
    return F.group_norm(x, groups, w.to(x.dtype), b.to(x.dtype), eps)


def _linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.to(x.dtype).t()
    return y + b.to(x.dtype) if b is not None else y


def _block(sd, pfx: str, spec: dict, x: Tensor, emb: Tensor) -> Tensor:
    """UNetBlock.forward (adm_blocks.py:159-181), adaptive_scale=True, dropout=0, skip_scale=1."""
    g = lambda n: sd.get(pfx + n)  # noqa: E731
    orig = x
    x = _conv(F.silu(_gn(x, g("norm0.weight"), g("norm0.bias"))), g("conv0.weight"), g("conv0.bias"),
              up=spec["up"], down=spec["down"])
    params = _linear(emb, g("affine.weight"), g("affine.bias")).unsqueeze(2).unsqueeze(3)
    scale, shift = params.chunk(2, dim=1)
    x = F.silu(torch.addcmul(shift, _gn(x, g("norm1.weight"), g("norm1.bias")), scale + 1))
    x = _conv(x, g("conv1.weight"), g("conv1.bias"))
    if spec["cin"] != spec["cout"] or spec["up"] or spec["down"]:
        x = x + _conv(orig, g("skip.weight"), g("skip.bias"), up=spec["up"], down=spec["down"])
    else:
        x = x + orig
    if spec["attn"]:
        n, c, h, w = x.shape
        heads = c // 64
        qkv = _conv(_gn(x, g("norm2.weight"), g("norm2.bias")), g("qkv.weight"), g("qkv.bias"))
        q, k, v = qkv.reshape(n * heads, c // heads, 3, -1).unbind(2)
        wts = torch.einsum("ncq,nck->nqk", q.float(), (k / math.sqrt(k.shape[1])).float()).softmax(dim=2).to(q.dtype)
        a = torch.einsum("nqk,nck->ncq", wts, v)
        x = _conv(a.reshape(n, c, h, w), g("proj.weight"), g("proj.bias")) + x
    return x


def unet_forward(sd: Dict[str, Tensor], model_cfg, x: Tensor, noise_labels: Tensor,
                 cond: Optional[Tensor] = None, x_self_cond: Optional[Tensor] = None) -> Tensor:
    """DhariwalUNet.forward (adm_blocks.py:364-404) for cat_cond=True, no labels / augment; self-conditioning stacks
    `x_self_cond` (zeros when None) in front of x (:321-324)."""
    ch = model_cfg["ch"]
    half = ch // 2
    freqs = torch.arange(half, device=noise_labels.device).to(noise_labels.dtype) / half          # PositionalEmbedding, :192-199
    freqs = (1 / 10000) ** freqs
    e = noise_labels.ger(freqs)
    emb = torch.cat([e.cos(), e.sin()], dim=1)
    emb = F.silu(_linear(emb, sd["map_layer0.weight"], sd["map_layer0.bias"]))
    emb = F.silu(_linear(emb, sd["map_layer1.weight"], sd["map_layer1.bias"]))
    if model_cfg.get("self_cond", False):
        x = torch.cat([torch.zeros_like(x) if x_self_cond is None else x_self_cond, x], dim=1)
    cc = model_cfg.get("cond_channels", 0) if model_cfg.get("cat_cond", False) else 0
    if cc > 0:
        if cond is None:
            cond = torch.zeros(x.shape[0], cc, x.shape[2], x.shape[3], dtype=x.dtype, device=x.device)
        x = torch.cat([cond, x], dim=1)                                 # :327-332, order [cond, x]
    plan = unet_plan(model_cfg)
    skips = []
    for spec in plan["enc"]:
        pfx = f"enc.{spec['name']}."
        if spec["kind"] == "conv":
            x = _conv(x, sd[pfx + "weight"], sd[pfx + "bias"])
        else:
            x = _block(sd, pfx, spec, x, emb)
        skips.append(x)
    for spec in plan["dec"]:
        if x.shape[1] != spec["cin"]:
            x = torch.cat([x, skips.pop()], dim=1)
        x = _block(sd, f"dec.{spec['name']}.", spec, x, emb)
    x = _conv(F.silu(_gn(x, sd["out_norm.weight"], sd["out_norm.bias"])), sd["out_conv.weight"], sd["out_conv.bias"])
    return x


# ------------------------------------------------------------------------------------------------
# EDM preconditioning / loss  (sigma_data = 1, mcedm.py:45-50)
# ------------------------------------------------------------------------------------------------
def precond_coeffs(sigma: Tensor):
    s = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    den = s ** 2 + 1.0
    return 1.0 / den, s / den.sqrt(), 1 / den.sqrt(), s.log() / 4   # c_skip, c_out, c_in, c_noise


def denoise(sd, model_cfg, xt: Tensor, sigma: Tensor, cond: Optional[Tensor]):
    """get_denoised / model_precond (mcedm.py:443-461, :199-211): returns (D_x, F_x) in fp32."""
    xt = xt.to(torch.float32)
    c_skip, c_out, c_in, c_noise = precond_coeffs(sigma)
    f_x = unet_forward(sd, model_cfg, c_in * xt, c_noise.flatten(), cond)
    return c_skip * xt + c_out * f_x, f_x


def loss_weight(sigma: Tensor) -> Tensor:
    return (sigma ** 2 + 1.0) / sigma ** 2                           # mcedm.py:237-239


def training_loss(sd, model_cfg, x: Tensor, sigma: Tensor, noise: Tensor, cond: Tensor, mask: Optional[Tensor]):
    """PlMcedm.forward + NoiseEstimationLoss (mcedm.py:213-235, :278; losses.py:48-53). NCHW inputs."""
    x_noise = x + mask * noise * sigma if mask is not None else x + noise * sigma
    d_x, _ = denoise(sd, model_cfg, x_noise, sigma, cond)
    w = loss_weight(sigma)
    if mask is not None:
        per = (w * (d_x * mask - x * mask) ** 2).sum(dim=(1, 2, 3))
    else:
        per = (w * (d_x - x) ** 2).sum(dim=(1, 2, 3))
    return per.mean(), d_x


def get_cond_in(x_bhwc: Tensor, mask_bhwc: Tensor, randn: Tensor) -> Tensor:
    """mcedm.py:241-252 without add_cond_mask/add_xt: observed values where mask==0, N(0,1) elsewhere."""
    return x_bhwc * (1 - mask_bhwc) + randn * mask_bhwc


def masked_mae(pred: Tensor, target: Tensor, mask: Tensor, loss_dim=None) -> Tensor:
    """MaskedLoss('l1') (losses.py:62-78)."""
    pred, target = pred * mask, target * mask
    if loss_dim is None:
        return (pred - target).abs().sum() / mask.sum()
    return (pred[..., loss_dim] - target[..., loss_dim]).abs().sum() / mask[..., loss_dim].sum()


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
def edm_schedule(num_steps: int, sigma_min: float, sigma_max: float, rho: float) -> Tensor:
    """rho-schedule in fp64 with t_N = 0 appended (mcedm.py:579-588)."""
    sigma_min = max(sigma_min, 0.002)
    sigma_max = min(sigma_max, 80)
    i = torch.arange(num_steps, dtype=torch.float64)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def sample_edm(sd, model_cfg, hu_noise: Tensor, cond: Tensor, hu_mask: Tensor, sparams,
               step_noise: Callable[[int, Tensor], Tensor], n_state_ch: Optional[int] = None,
               return_last: bool = True, record: Optional[list] = None) -> Tensor:
    """PlMcedm.sample_edm (mcedm.py:570-638) with guidance off (guide_dx False, dx_cond False, w = 0).

    hu_noise  fp32 [B,C,H,W]: the draw of `torch.randn_like(hu)` (:576)
    step_noise(i, x_cur) -> fp64 tensor: the draw of `torch.randn_like(x_cur)` of step i (:608)
    record: optional list receiving (step, which, sigma, D_x) for per-step denoiser parity checks.
    Returns xs [B, T, H, W, C] fp64 (T = 1 when return_last).
    """
    num_steps = int(sparams["timesteps"])
    t_steps = edm_schedule(num_steps, sparams["sigma_min"], sparams["sigma_max"], sparams["rho"])
    c = n_state_ch if n_state_ch is not None else hu_noise.shape[1]
    known = cond[:, 0:c]
    x_next = hu_noise.to(torch.float64) * t_steps[0]
    x_next = known * (1 - hu_mask) + x_next * hu_mask
    xs = [x_next]
    s_min, s_max = sparams["S_min"], float(sparams["S_max"])
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        gamma = min(sparams["S_churn"] / num_steps, np.sqrt(2) - 1) if s_min <= t_cur <= s_max else 0
        t_hat = t_cur + gamma * t_cur
        x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * sparams["S_noise"] * step_noise(i, x_cur) * hu_mask
        d1, _ = denoise(sd, model_cfg, x_hat, t_hat, cond)
        if record is not None:
            record.append((i, 0, float(t_hat), d1))
        d_cur = (x_hat - d1.to(torch.float64)) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur * hu_mask
        if i < num_steps - 1:
            d2, _ = denoise(sd, model_cfg, x_next, t_next, cond)
            if record is not None:
                record.append((i, 1, float(t_next), d2))
            d_prime = (x_next - d2.to(torch.float64)) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime) * hu_mask
        xs = [x_next] if return_last else xs + [x_next]
    return torch.stack(xs, dim=0).permute(1, 0, 3, 4, 2)             # 't b c h w -> b t h w c'


def cond_sample_edm(sd, model_cfg, u_noise: Tensor, h_cond: Tensor, sparams,
                    step_noise: Callable[[int, Tensor], Tensor], return_last: bool = True,
                    record: Optional[list] = None,
                    guide: Optional[Callable[[Tensor, Tensor], Tensor]] = None) -> Tensor:
    """PlCondDdim.sample_edm as PlCondEdm uses it (models/ddim.py:1532-1601; config 5): plain stochastic Heun, NO
    mask, the condition is h; w = 0; no self-conditioning.  u_noise fp32 [B,C,H,W] is the caller's draw (the sampler
    does not draw its own initial noise, unlike PlMcedm); h_cond [B,Cc,H,W].  Returns xs [B, T, H, W, C] fp64.
    guide(h_cond, denoised fp64) -> fp32 [B,1,H,W]: `get_dx_log_prob(h, denoised, guide_dx=True)` (ddim.py:641-650,
    see oracle/pde_oracle.get_dx_pde_cond); None => guide_dx False (the `- 5 * dx / t_hat` terms are exact zeros)."""
    num_steps = int(sparams["timesteps"])
    t_steps = edm_schedule(num_steps, sparams["sigma_min"], sparams["sigma_max"], sparams["rho"])
    x_next = u_noise.to(torch.float64) * t_steps[0]                   # :1555
    xs = [x_next]
    s_min, s_max = sparams["S_min"], float(sparams["S_max"])
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        gamma = min(sparams["S_churn"] / num_steps, np.sqrt(2) - 1) if s_min <= t_cur <= s_max else 0
        t_hat = t_cur + gamma * t_cur
        x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * sparams["S_noise"] * step_noise(i, x_cur)   # :1566
        d1, _ = denoise(sd, model_cfg, x_hat, t_hat, h_cond)
        if record is not None:
            record.append((i, 0, float(t_hat), d1))
        d1 = d1.to(torch.float64)
        d_cur = (x_hat - d1) / t_hat                                  # :1569-1571
        if guide is not None:
            d_cur = d_cur - 5. * guide(h_cond, d1) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            d2, _ = denoise(sd, model_cfg, x_next, t_next, h_cond)
            if record is not None:
                record.append((i, 1, float(t_next), d2))
            d2 = d2.to(torch.float64)
            d_prime = (x_next - d2) / t_next                          # :1588-1590 (t_hat under dx, as the reference)
            if guide is not None:
                d_prime = d_prime - 5. * guide(h_cond, d2) / t_hat
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
        xs = [x_next] if return_last else xs + [x_next]
    return torch.stack(xs, dim=0).permute(1, 0, 3, 4, 2)


def cond_training_loss(sd, model_cfg, u: Tensor, sigma: Tensor, noise: Tensor, h_cond: Optional[Tensor]):
    """PlCondEdm.forward + NoiseEstimationLoss (models/ddim.py:1668-1694, :1723; losses.py:48-53): x_noise = u + noise*sigma,
    loss over every pixel, the condition h is not scaled (and is None when the cond_p coin drops it)."""
    return training_loss(sd, model_cfg, u, sigma, noise, h_cond, None)


# ------------------------------------------------------------------------------------------------
# PlDdim: EDM sampler on the VP sigma grid with RePaint-style conditioning (models/ddim.py:915-1051), BASELINE config 4
# ------------------------------------------------------------------------------------------------
class VpGrid:
    """The diffusion schedule of PlDdim (ddim.py:150-160, ddim_blocks.py:487-490 linear betas) and what
    set_test_sampler_params derives from it (:121-137); round_sigma (:949-957) and compute_alpha (:700-704)."""

    def __init__(self, beta_start=0.0001, beta_end=0.02, n=1000):
        self.betas = torch.from_numpy(np.linspace(beta_start, beta_end, n, dtype=np.float64)).float()
        self.num_timesteps = n
        alphas_bar = (1.0 - self.betas).cumprod(dim=0)
        self.edm_steps = ((1 - alphas_bar) / alphas_bar).sqrt().flip(dims=(0,))
        self.sigma_min = float(self.edm_steps[n - 1])
        self.sigma_max = float(self.edm_steps[0])

    def round_sigma(self, sigma: Tensor, return_index=False) -> Tensor:
        sigma32 = sigma.to(torch.float32)
        index = torch.cdist(sigma32.reshape(1, -1, 1), self.edm_steps.reshape(1, -1, 1)).argmin(2)
        result = index if return_index else self.edm_steps[index.flatten()]
        return result.type_as(sigma).reshape(sigma.shape)

    def compute_alpha(self, t: Tensor) -> Tensor:
        betas = torch.cat([torch.zeros(1), self.betas], dim=0)
        return (1 - betas).cumprod(dim=0).index_select(0, t.reshape(-1) + 1).view(-1, 1, 1, 1)


def vp_denoise(sd, model_cfg, grid: VpGrid, xt: Tensor, t: Tensor, net=None):
    """PlDdim.get_denoised with cond = x_self_cond = dx = None (ddim.py:915-947): c_skip = 1, c_out = -sigma.
    net(sd, model_cfg, x, noise_labels): the network — `unet_forward` (ADM branch, `hparams.name` = adm*) by default,
    `oracle.ddpm_oracle.ddpm_unet_forward` for the DDPM U-Net branch (ddim.py:40-43)."""
    xt = xt.to(torch.float32)
    sigma = t.to(torch.float32).reshape(-1, 1, 1, 1)
    c_in = 1 / (sigma ** 2 + 1).sqrt()
    c_noise = grid.num_timesteps - 1 - grid.round_sigma(sigma, return_index=True).to(torch.float32)
    f_x = (net or unet_forward)(sd, model_cfg, c_in * xt, c_noise.flatten())
    return 1 * xt + (-sigma) * f_x, f_x


def ddim_sample_edm(sd, model_cfg, grid: VpGrid, hu: Tensor, hu_noise: Tensor, sparams,
                    step_noise: Callable[[Tensor], Tensor], h_ch: int = 1, u_ch: int = 1, return_last: bool = True,
                    record: Optional[list] = None, net=None) -> Tensor:
    """PlDdim.sample_edm (ddim.py:959-1051) with guide_dx False, w = 0.  hu fp32 [B,C,H,W]: the normalised ground truth
    whose first n_time_h / n_time_u time rows are known (mask == 1 KNOWN); hu_noise: the draw of :967;
    step_noise(x) -> fp64 draw of :1002 / :1036.  Returns xs [B, T, H, W, C] fp64."""
    n_repeat, n_time_h, n_time_u = sparams["n_repeat"], sparams["n_time_h"], sparams["n_time_u"]
    hu_mask = torch.ones_like(hu)
    hu_mask[:, 0:h_ch, n_time_h:, :] = 0.0
    hu_mask[:, h_ch:h_ch + u_ch, n_time_u:, :] = 0.0
    sigma_min = max(sparams["sigma_min"], grid.sigma_min)
    sigma_max = min(sparams["sigma_max"], grid.sigma_max)
    num_steps, rho = int(sparams["timesteps"]), sparams["rho"]
    i_ = torch.arange(num_steps, dtype=torch.float64)
    t_steps = (sigma_max ** (1 / rho) + i_ / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    t_steps = torch.cat([grid.round_sigma(t_steps), torch.zeros_like(t_steps[:1])])
    aT = grid.compute_alpha(t_steps[0].long())
    hu_t_known = hu * aT.sqrt() + hu_noise * (1.0 - aT).sqrt()
    x = hu_t_known * hu_mask + hu_noise * (1.0 - hu_mask)
    x_next = x.to(torch.float64) * t_steps[0]
    xs = [x_next]
    s_min, s_max = sparams["S_min"], float(sparams["S_max"])
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        gamma = min(sparams["S_churn"] / num_steps, np.sqrt(2) - 1) if s_min <= t_cur <= s_max else 0
        t_hat = grid.round_sigma(t_cur + gamma * t_cur)
        x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * sparams["S_noise"] * step_noise(x_cur)
        for k in range(n_repeat):
            d1, _ = vp_denoise(sd, model_cfg, grid, x_hat, t_hat, net)
            if record is not None:
                record.append((i, k, 0, float(t_hat), d1))
            d_cur = (x_hat - d1.to(torch.float64)) / t_hat
            x_next = x_hat + (t_next - t_hat) * d_cur
            if i < num_steps - 1:
                d2, _ = vp_denoise(sd, model_cfg, grid, x_next, t_next, net)
                if record is not None:
                    record.append((i, k, 1, float(t_next), d2))
                d_prime = (x_next - d2.to(torch.float64)) / t_next
                x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
            at_next = grid.compute_alpha(t_next.long())
            hu_t_known = at_next.sqrt() * hu + (1 - at_next).sqrt() * hu_noise
            x_next = hu_t_known * hu_mask + x_next * (1.0 - hu_mask)
            if k < n_repeat - 1:
                t_hat = grid.round_sigma(t_next + (np.sqrt(2) - 1) * t_next)
                x_hat = x_next + (t_hat ** 2 - t_next ** 2).sqrt() * sparams["S_noise"] * step_noise(x_next)
        if i == num_steps - 1:
            x_next = hu * hu_mask + x_next * (1.0 - hu_mask)
        xs = [x_next] if return_last else xs + [x_next]
    return torch.stack(xs, dim=0).permute(1, 0, 3, 4, 2)


def ddim_sample_with_repeat(sd, model_cfg, grid: VpGrid, hu: Tensor, hu_noise: Tensor, sparams, net, h_ch: int = 1,
                            u_ch: int = 1, return_last: bool = True, record: Optional[list] = None,
                            rand_like: Optional[Callable[[Tensor], Tensor]] = None):
    """PlDdim.sample_with_repeat (ddim.py:808-913) with guide_dx False, w = 0, dx_cond False: DDIM steps on the VP
    schedule, the known region (mask == 1) re-imposed on every x0 prediction and on every x_t, n_repeat evaluations per
    timestep, self-conditioning on the previous x0 prediction.  hu fp32 [B,C,H,W]; hu_noise: the draw of :836;
    net(sd, model_cfg, x, t, x_self_cond) -> e_t.  Returns (xs, x0_preds) as [B, T, H, W, C] fp32."""
    n_repeat, n_time_h, n_time_u = sparams["n_repeat"], sparams["n_time_h"], sparams["n_time_u"]
    hu_mask = torch.ones_like(hu)
    hu_mask[:, 0:h_ch, n_time_h:, :] = 0.0
    hu_mask[:, h_ch:h_ch + u_ch, n_time_u:, :] = 0.0
    T = grid.num_timesteps
    if sparams["skip_type"] == "uniform":
        seq = list(range(0, T, T // int(sparams["timesteps"])))
    elif sparams["skip_type"] == "quad":
        seq = [int(v) for v in (np.linspace(0, np.sqrt(T * 0.8), int(sparams["timesteps"])) ** 2)]
    else:
        raise NotImplementedError
    a = (1 - grid.betas).cumprod(dim=0)
    hu_t_known = hu * a[T - 1].sqrt() + hu_noise * (1.0 - a[T - 1]).sqrt()
    x = hu_t_known * hu_mask + hu_noise * (1.0 - hu_mask)
    n = hu.size(0)
    seq_next = [-1] + list(seq[:-1])
    xs, x0_preds, x0_t = [x], [], None
    self_cond = bool(model_cfg.get("self_cond", False))
    for i, j in zip(reversed(seq), reversed(seq_next)):
        t = (torch.ones(n) * i).type_as(hu)
        next_t = (torch.ones(n) * j).type_as(hu)
        at, at_next = grid.compute_alpha(t.long()), grid.compute_alpha(next_t.long())
        xt = xs[-1]
        for k in range(n_repeat):
            et = net(sd, model_cfg, xt, t, x0_t if self_cond else None)
            if record is not None:
                record.append(dict(t=float(i), k=k, xt=xt, x_self_cond=x0_t if self_cond else None, et=et))
            et = et - 5.0 * (1 - at).sqrt() * torch.zeros_like(xt)
            x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
            x0_t = hu * hu_mask + x0_t * (1.0 - hu_mask)
            if k < n_repeat - 1:
                xt = at.sqrt() * x0_t + (1 - at).sqrt() * et
        if abs(sparams["eta"]) > 1e-10:
            c1 = sparams["eta"] * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
            c2 = ((1 - at_next) - c1 ** 2).sqrt()
            xt_next = at_next.sqrt() * x0_t + c1 * rand_like(x) + c2 * et
        else:
            c2 = (1 - at_next).sqrt()
            xt_next = at_next.sqrt() * x0_t + c2 * et
        hu_t_known = at_next.sqrt() * hu + c2 * hu_noise
        xt_next = hu_t_known * hu_mask + xt_next * (1.0 - hu_mask)
        if return_last:
            x0_preds, xs = [x0_t], [xt_next]
        else:
            x0_preds.append(x0_t)
            xs.append(xt_next)
    return torch.stack(xs, dim=0).permute(1, 0, 3, 4, 2), torch.stack(x0_preds, dim=0).permute(1, 0, 3, 4, 2)


# ------------------------------------------------------------------------------------------------
# mask generators (datamodules/h5_dataset.py), mask == 1 -> missing / to be generated
# ------------------------------------------------------------------------------------------------
def train_mask(inp: Tensor, target: Tensor, coin: float) -> Tensor:
    """HDF5MaskDataset.sample_mask, is_train (h5_dataset.py:235-243): coin = torch.rand(1) draw."""
    if coin > 0.5:
        return torch.cat([torch.zeros_like(inp), torch.ones_like(target)], dim=-1)
    return torch.cat([torch.ones_like(inp), torch.zeros_like(target)], dim=-1)


def eval_masks(inp: Tensor, target: Tensor) -> Dict[str, Tensor]:
    """HDF5MaskDataset.sample_mask, eval (h5_dataset.py:245-253): 'u' = u missing, 'h' = h missing."""
    return {"u": torch.cat([torch.zeros_like(inp), torch.ones_like(target)], dim=-1),
            "h": torch.cat([torch.ones_like(inp), torch.zeros_like(target)], dim=-1)}


def time_train_mask(inp: Tensor, target: Tensor, var: float, t_max1: int, t_max2: int) -> Tensor:
    """HDF5TimeMaskDataset.get_train_mask (h5_dataset.py:306-337); var, t_max* are the RNG draws."""
    ic = inp.shape[-1]
    if var <= 0.4:
        mv = torch.cat([torch.zeros_like(inp, dtype=torch.bool), torch.ones_like(target, dtype=torch.bool)], -1)
    elif var <= 0.8:
        mv = torch.cat([torch.ones_like(inp, dtype=torch.bool), torch.zeros_like(target, dtype=torch.bool)], -1)
    else:
        mv = torch.cat([torch.zeros_like(inp, dtype=torch.bool), torch.zeros_like(target, dtype=torch.bool)], -1)
    mr = torch.ones_like(mv, dtype=torch.bool)
    mr[:t_max1, :, :ic] = False
    mr[:t_max2, :, ic:] = False
    return (mv | mr).float()


def time_eval_masks(inp: Tensor, target: Tensor) -> Dict[str, Tensor]:
    """HDF5TimeMaskDataset.sample_mask, eval with add_time_masks (h5_dataset.py:356-391)."""
    half = int(0.5 * inp.shape[0])
    zi, zt, oi, ot = torch.zeros_like(inp), torch.zeros_like(target), torch.ones_like(inp), torch.ones_like(target)
    a, b = zi.clone(), zt.clone()
    a[half:] = 1
    b[half:] = 1
    m_hu = torch.cat([a, b], -1)
    a = zi.clone()
    a[half:] = 1
    m_u = torch.cat([a, ot], -1)
    b = zt.clone()
    b[half:] = 1
    m_h = torch.cat([oi, b], -1)
    return {"hu": m_hu, "u": m_u, "h": m_h}
