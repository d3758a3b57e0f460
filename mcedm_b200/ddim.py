"""Host mirror of the reference's joint-state baseline module `PlDdim` (models/ddim.py:17-1051) for its EDM sampler with
RePaint-style conditioning — BASELINE config 4 (`diff_sampler=edm_sampler, n_time_h=0, n_time_u=64`):

    get_diffusion_schedule / get_edm_steps / set_test_sampler_params     models/ddim.py:116-137, :150-160
    compute_alpha / round_sigma                                           :700-704, :949-957
    get_denoised (VP preconditioning: c_skip = 1, c_out = -sigma, c_in = 1/sqrt(sigma^2+1),
                  c_noise = T-1-index(sigma))                              :915-947
    sample_edm (stochastic Heun on the VP sigma grid; the known region is re-imposed at the noise level of every
                step and the step is repeated n_repeat times after re-noising)   :959-1051
    test_step (MAE / masked MAE / known-region metrics)                    :372-533 (core metrics)

The reference builds the network from `hparams.name`: `DhariwalUNet` when it starts with "adm", the DDPM U-Net
`ddim_blocks.Model` otherwise (:40-43).  Both run on the kernels: `configs/config_adm_ddim_res32.yaml` (ADM branch) and
the shipped `configs/config_ddim_res32.yaml` (`ddpm_blocks.Model` on `ddpm_engine.DdpmEngine`: stride-2 convs, 32-group
GroupNorm, 256-wide time embedding).

Kernel path: the Heun updates are the PlMcedm kernels with an all-ones mask (exact: `*1.0`), `c_skip = 1`,
`c_out = -sigma`; the known-region handling is `mcedm_edm_vp_init` / `mcedm_edm_repaint_blend` (mask == 1 means KNOWN
here — the opposite of PlMcedm).  Scalars (sigma grid look-ups, alpha-bar values) are evaluated on the host with the
reference's own torch expressions on CPU tensors.  Mask polarity, `t.long()` truncation of a *sigma* used as a timestep
index (:989, :1029) and the `t_hat` re-noising rule (:1035) are reproduced as they are.

    sample_with_repeat (DDIM steps on the VP schedule, known region re-imposed on every x0 prediction and every x_t,
                n_repeat evaluations per timestep, self-conditioning on the previous x0 prediction)   :808-913

Not mirrored: the DDPM noise-prediction training loss, the plain DDIM sampler `sample` (no conditioning), PDE guidance
(`guide_dx`) and `dx_cond` for this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from einops import rearrange

from . import _lib as L
from .mcedm import PlMcedm
from .nn_misc import CorrelationLoss, MaskedLoss


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):   # ddim_blocks.py:473-505
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_diffusion_timesteps, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(num_diffusion_timesteps, 1, num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        betas = np.linspace(-6, 6, num_diffusion_timesteps)
        betas = 1 / (np.exp(-betas) + 1) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    return torch.from_numpy(betas).float()


def _f32(x) -> np.float32:
    return np.float32(x)


class PlDdim(PlMcedm):
    def __init__(self, hparams):
        super().__init__(hparams)                                     # 'adm*' -> DhariwalUNet, else ddim_blocks.Model (:40-43)
        self.cond_p = 0.0                                             # :30
        self.model_var_type = hparams.model.var_type
        betas, posterior_variance = self.get_diffusion_schedule(hparams)
        self.register_buffer("betas", betas)
        self.num_timesteps = betas.shape[0]
        if self.model_var_type == "fixedlarge":
            self.register_buffer("logvar", betas.log())
        elif self.model_var_type == "fixedsmall":
            self.register_buffer("logvar", posterior_variance.clamp(min=1e-20).log())
        self.mae_criterion = torch.nn.L1Loss()                        # :75-77
        self.mae_criterion_mask = MaskedLoss()
        self.correlation = CorrelationLoss()
        self.edm_steps = None
        self._alphas_host = None

    @staticmethod
    def get_diffusion_schedule(hparams):                              # :150-160
        d = hparams.diffusion
        betas = get_beta_schedule(beta_schedule=d.beta_schedule, beta_start=d.beta_start, beta_end=d.beta_end,
                                  num_diffusion_timesteps=d.num_diffusion_timesteps)
        alphas_cumprod = (1.0 - betas).cumprod(dim=0)
        alphas_cumprod_prev = torch.cat([torch.ones(1), alphas_cumprod[:-1]], dim=0)
        posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
        return betas, posterior_variance

    def get_edm_steps(self):                                          # :131-137 (host copy: the grid is 1000 scalars)
        alphas_bar = (1.0 - self.betas.detach().cpu()).cumprod(dim=0)
        return ((1 - alphas_bar) / alphas_bar).sqrt().flip(dims=(0,))

    def set_test_sampler_params(self, params):                        # :121-129
        self.test_sparams = params
        if params.type == "edm":
            self.edm_steps = self.get_edm_steps()
            self.sigma_min = float(self.edm_steps[self.num_timesteps - 1])
            self.sigma_max = float(self.edm_steps[0])

    def compute_alpha(self, t):                                       # :700-704, on the host (t: int64 CPU tensor)
        if self._alphas_host is None:
            b = self.betas.detach().cpu()
            self._alphas_host = (1 - torch.cat([torch.zeros(1).type_as(b), b], dim=0)).cumprod(dim=0)
        return self._alphas_host.index_select(0, torch.as_tensor(t).reshape(-1) + 1).view(-1, 1, 1, 1)

    def round_sigma(self, sigma, return_index=False):                 # :949-957, CPU tensors
        if self.edm_steps is None:
            raise RuntimeError("call set_test_sampler_params(cfg.diff_sampler) first (run.py:87)")
        sigma = torch.as_tensor(sigma).detach().cpu()
        sigma32 = sigma.to(torch.float32)
        steps = self.edm_steps
        index = torch.cdist(sigma32.reshape(1, -1, 1), steps.reshape(1, -1, 1)).argmin(2)
        result = index if return_index else steps[index.flatten()]
        return result.type_as(sigma).reshape(sigma.shape)

    def _vp_scalars(self, sigma: float):
        """fp32 (c_out, c_in, c_noise) of one grid sigma with the op sequence of :921-925."""
        s = _f32(sigma)
        c_in = _f32(_f32(1.0) / np.sqrt(_f32(_f32(s * s) + _f32(1.0)), dtype=np.float32))
        idx = int(self.round_sigma(torch.tensor(float(s), dtype=torch.float32), return_index=True))
        return float(-s), float(c_in), float(_f32(self.num_timesteps - 1 - idx))

    # ---------------------------------------------------------------- network + preconditioning
    def get_denoised(self, model, xt, t, cond=None, x_self_cond=None, dx=None, w=None):   # :915-947
        if dx is not None:
            raise NotImplementedError("dx conditioning is not supported")
        if not (w is None or abs(w) < 0.001 or cond is None):
            raise NotImplementedError("classifier-free guidance (w != 0) is not supported")
        lib = L.lib()
        unet = self._unet_of(model)
        xt = xt.to(torch.float32).contiguous()
        B, chw = xt.shape[0], xt[0].numel()
        c_out, c_in, c_noise = self._vp_scalars(float(torch.as_tensor(t).reshape(-1)[0]))
        dev = xt.device
        ci = torch.tensor([c_in], device=dev)
        st = L.stream_ptr()
        x_in = torch.empty_like(xt)
        L.check(lib.mcedm_edm_precond_in(L.ptr(xt), L.ptr(ci), 0, B, chw, L.ptr(x_in), st), "precond_in")
        if x_self_cond is not None:
            x_self_cond = (ci.reshape(1, 1, 1, 1) * x_self_cond).to(torch.float32)
        if cond is not None and unet.cat_condition:
            cond = (ci.reshape(1, 1, 1, 1) * cond).to(torch.float32)       # :931-932: a concatenated cond is scaled too
        F_x = unet(x_in, torch.tensor([c_noise], device=dev), cond, x_self_cond=x_self_cond)
        D_x = torch.empty_like(xt)
        coef = torch.tensor([1.0, c_out], device=dev)                 # c_skip = 1, c_out = -sigma (kept alive: L.ptr)
        L.check(lib.mcedm_edm_precond_out(L.ptr(xt), L.ptr(F_x), L.ptr(coef), L.ptr(coef[1:]), 0, B, chw, L.ptr(D_x),
                                          st), "precond_out")
        return D_x, F_x

    def training_step(self, train_batch, batch_idx):
        raise NotImplementedError("the DDPM noise-prediction training loss of PlDdim is outside the hot path")

    # ---------------------------------------------------------------- sampler
    def _edm_grid(self, sparams):
        """t_steps of :975-984 on the host: rho-schedule snapped to the VP sigma grid, t_N = 0."""
        sigma_min = max(sparams.sigma_min, self.sigma_min)
        sigma_max = min(sparams.sigma_max, self.sigma_max)
        n = sparams.timesteps
        i = torch.arange(n, dtype=torch.float64)
        t = (sigma_max ** (1 / sparams.rho) + i / (n - 1) * (sigma_min ** (1 / sparams.rho)
                                                            - sigma_max ** (1 / sparams.rho))) ** sparams.rho
        return torch.cat([self.round_sigma(t), torch.zeros_like(t[:1])])

    def _known_coeffs(self, t):
        """(sqrt(a), sqrt(1-a)) in fp32 with a = compute_alpha(t.long())  (:989-990, :1029-1030)."""
        a = self.compute_alpha(torch.as_tensor(t).long())
        return float(a.sqrt().reshape(())), float((1.0 - a).sqrt().reshape(()))

    @torch.no_grad()
    def sample_edm(self, h, u, sparams, return_last=True, guide_dx=False):
        """h, u: normalised ground-truth channels `b h w c` (the first n_time_h / n_time_u time rows are the observed
        part). Returns xs [b, t, h, w, c] float64."""
        if guide_dx:
            raise NotImplementedError("guide_dx is not supported by PlDdim on the kernels")
        w = sparams.w
        if not (w is None or abs(w) < 0.001):
            raise NotImplementedError("classifier-free guidance (w != 0) is not supported")
        if not h.is_cuda:
            raise L.McedmError("sample_edm needs CUDA tensors: the sm_100a kernels have no CPU fallback")
        lib = L.lib()
        n_repeat, n_time_h, n_time_u = sparams.n_repeat, sparams.n_time_h, sparams.n_time_u
        model = self.ema_model if self.ema_model is not None else self.model
        unet = self._unet_of(model)
        hu = rearrange(torch.cat([h, u], dim=-1), "b h w c -> b c h w").contiguous().float()
        hu_noise = self._randn_like("init", hu)                       # :967
        hu_mask = torch.ones_like(hu)
        hu_mask[:, 0:self.h_ch, n_time_h:, :] = 0.0
        hu_mask[:, self.h_ch:self.h_ch + self.u_ch, n_time_u:, :] = 0.0
        B, C, H, W = hu.shape
        total, dev = hu.numel(), hu.device
        num_steps = sparams.timesteps
        t_steps = self._edm_grid(sparams)
        bufs = self._sampler_buffers(B, C, H, W, max(unet.cat_channels, 1), unet.out_channels, dev)
        x_cur, x_hat, x_e, d_cur = bufs["x_cur"], bufs["x_hat"], bufs["x_e"], bufs["d_cur"]
        x_in, F_buf, nl = bufs["x_in"], bufs["F"], bufs["nl"]
        ones = bufs["mask"]
        ones.fill_(1.0)
        cat_in = None
        if unet.cat_channels > 0:                                     # cond = None and x_self_cond = None: zeros (:324, :329)
            cat_in = bufs["cond"]
            cat_in.zero_()
        D_buf = torch.empty_like(x_in) if self._trace is not None else None
        st = L.stream_ptr()
        engine = unet.engine()
        x_in_c, cat_c = engine._check_inputs(x_in, nl, cat_in)[0::2]
        assert x_in_c.data_ptr() == x_in.data_ptr()

        def net_eval(c_noise):
            nl.fill_(c_noise)
            return engine.forward_static(x_in, nl, cat_c, F_buf, use_graph=self.use_cuda_graph)

        sa, s1 = self._known_coeffs(t_steps[0])
        hu_noise = hu_noise.to(torch.float32).contiguous()
        L.check(lib.mcedm_edm_vp_init(L.ptr(hu), L.ptr(hu_noise), L.ptr(hu_mask), sa, s1, float(t_steps[0]), total,
                                      L.ptr(x_cur), st), "edm_vp_init")
        xs = [x_cur.clone()] if not return_last else None
        S_min, S_max = sparams.S_min, float(sparams.S_max)
        for i in range(num_steps):
            t_cur, t_next = t_steps[i], t_steps[i + 1]
            gamma = min(sparams.S_churn / num_steps, np.sqrt(2) - 1) if S_min <= t_cur <= S_max else 0
            t_hat = self.round_sigma(t_cur + gamma * t_cur)
            last = i == num_steps - 1
            src = x_cur
            t_from = t_cur
            for k in range(n_repeat):
                # x_hat = src + sqrt(t_hat^2 - t_from^2) * S_noise * randn   (:1002, :1036)
                coef = float((t_hat ** 2 - t_from ** 2).sqrt() * sparams.S_noise)
                eps = self._randn_like("step", src)
                c_out, c_in, c_noise = self._vp_scalars(float(t_hat))
                L.check(lib.mcedm_edm_churn(L.ptr(src), L.ptr(eps), L.ptr(ones), coef, c_in, total, L.ptr(x_hat),
                                            L.ptr(x_in), st), "edm_churn")
                F1 = net_eval(c_noise)
                c_out2, c_in2, c_noise2 = self._vp_scalars(float(t_next)) if not last else (0.0, 0.0, 0.0)
                out_e = x_cur if last else x_e
                L.check(lib.mcedm_edm_euler(L.ptr(x_hat), L.ptr(F1), L.ptr(ones), float(t_hat), float(t_next), 1.0,
                                            c_out, c_in2, total, L.ptr(d_cur), L.ptr(out_e),
                                            None if last else L.ptr(x_in), L.ptr(D_buf), st), "edm_euler")
                if self._trace is not None:
                    self._trace.append((i, k, 0, float(t_hat), D_buf.clone(), x_hat.clone()))
                if not last:
                    F2 = net_eval(c_noise2)
                    L.check(lib.mcedm_edm_correct(L.ptr(x_hat), L.ptr(x_e), L.ptr(F2), L.ptr(d_cur), L.ptr(ones),
                                                  float(t_hat), float(t_next), 1.0, c_out2, total, L.ptr(x_cur),
                                                  L.ptr(D_buf), st), "edm_correct")
                    if self._trace is not None:
                        self._trace.append((i, k, 1, float(t_next), D_buf.clone(), x_e.clone()))
                sa, s1 = self._known_coeffs(t_next)                   # replace the known part (:1029-1031)
                L.check(lib.mcedm_edm_repaint_blend(L.ptr(hu), L.ptr(hu_noise), L.ptr(hu_mask), sa, s1, total,
                                                    L.ptr(x_cur), st), "edm_repaint_blend")
                if k < n_repeat - 1:                                  # back from t_next to a new t_hat (:1033-1036)
                    gamma1 = np.sqrt(2) - 1
                    t_hat = self.round_sigma(t_next + gamma1 * t_next)
                    src, t_from = x_cur, t_next
            if last:                                                  # :1038-1041
                L.check(lib.mcedm_edm_repaint_blend(L.ptr(hu), L.ptr(hu_noise), L.ptr(hu_mask), 1.0, 0.0, total,
                                                    L.ptr(x_cur), st), "edm_repaint_blend")
            if xs is not None:
                xs.append(x_cur.clone())
        xs = torch.stack(xs, dim=0) if xs is not None else x_cur.clone().unsqueeze(0)
        return rearrange(xs, "t b c h w -> b t h w c")

    @torch.no_grad()
    def sample_with_repeat(self, h, u, sparams, return_last=True, guide_dx=False):
        """PlDdim.sample_with_repeat (:808-913) on the kernels: returns (xs, x0_preds), each [b, t, h, w, c] float32.
        The state updates are mcedm_ddim_init / _x0 / _next (bit-identical to the torch expressions, scalars evaluated on
        the host with the reference's fp32 tensor ops); every e_t is one evaluation of the network with the per-sample
        timestep vector and the previous x0 prediction as self-conditioning input."""
        if guide_dx:
            raise NotImplementedError("guide_dx is not supported by PlDdim on the kernels")
        w = sparams.w
        if not (w is None or abs(w) < 0.001):
            raise NotImplementedError("classifier-free guidance (w != 0) is not supported")
        if not h.is_cuda:
            raise L.McedmError("sample_with_repeat needs CUDA tensors: the sm_100a kernels have no CPU fallback")
        lib = L.lib()
        model = self.ema_model if self.ema_model is not None else self.model
        unet = self._unet_of(model)
        n_repeat, n_time_h, n_time_u = sparams.n_repeat, sparams.n_time_h, sparams.n_time_u
        hu = rearrange(torch.cat([h, u], dim=-1), "b h w c -> b c h w").contiguous().float()
        hu_mask = torch.ones_like(hu)
        hu_mask[:, 0:self.h_ch, n_time_h:, :] = 0.0
        hu_mask[:, self.h_ch:self.h_ch + self.u_ch, n_time_u:, :] = 0.0
        T = self.num_timesteps
        if sparams.skip_type == "uniform":
            seq = list(range(0, T, T // sparams.timesteps))
        elif sparams.skip_type == "quad":
            seq = [int(v) for v in (np.linspace(0, np.sqrt(T * 0.8), sparams.timesteps) ** 2)]
        else:
            raise NotImplementedError
        hu_noise = self._randn_like("init", hu).to(torch.float32).contiguous()      # :836
        a = (1 - self.betas.detach().cpu()).cumprod(dim=0)                          # :838, host copy
        total, st, n = hu.numel(), L.stream_ptr(), hu.shape[0]
        x = torch.empty_like(hu)
        L.check(lib.mcedm_ddim_init(L.ptr(hu), L.ptr(hu_noise), L.ptr(hu_mask), float(a[T - 1].sqrt()),
                                    float((1.0 - a[T - 1]).sqrt()), total, L.ptr(x), st), "ddim_init")
        seq_next = [-1] + list(seq[:-1])
        xs, x0_preds, x0_t = [x], [], None
        self_cond = bool(getattr(unet, "self_condition", False))
        for i, j in zip(reversed(seq), reversed(seq_next)):
            t_host = torch.ones(n) * i
            at = self.compute_alpha(t_host.long())[0].reshape(())                   # fp32 scalars as the reference forms them
            at_next = self.compute_alpha((torch.ones(n) * j).long())[0].reshape(())
            sa, s1 = float(at.sqrt()), float((1 - at).sqrt())
            t_dev = t_host.to(hu.device, torch.float32)
            xt = xs[-1]
            et = None
            for k in range(n_repeat):
                et = unet(xt, t_dev, x_self_cond=x0_t if self_cond else None).contiguous()
                if self._trace is not None:
                    self._trace.append(dict(t=float(i), k=k, xt=xt.clone(), x_self_cond=None if x0_t is None else x0_t.clone(),
                                            et=et.clone()))
                x0_new = torch.empty_like(hu)
                xt_new = torch.empty_like(hu) if k < n_repeat - 1 else None
                L.check(lib.mcedm_ddim_x0(L.ptr(xt), L.ptr(et), L.ptr(hu), L.ptr(hu_mask), sa, s1, total, L.ptr(x0_new),
                                          L.ptr(xt_new), st), "ddim_x0")
                x0_t = x0_new
                if xt_new is not None:
                    xt = xt_new
            rnd, c1 = None, 0.0
            if abs(sparams.eta) > 1e-10:
                c1_t = sparams.eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
                c2_t = ((1 - at_next) - c1_t ** 2).sqrt()
                c1, c2 = float(c1_t), float(c2_t)
                rnd = (self._noise_hook("rand", x) if self._noise_hook is not None else torch.rand_like(x)).contiguous()
            else:
                c2 = float((1 - at_next).sqrt())
            x_next = torch.empty_like(hu)
            L.check(lib.mcedm_ddim_next(L.ptr(x0_t), L.ptr(et), L.ptr(hu), L.ptr(hu_noise), L.ptr(hu_mask), L.ptr(rnd),
                                        float(at_next.sqrt()), c1, c2, total, L.ptr(x_next), st), "ddim_next")
            if return_last:
                x0_preds, xs = [x0_t], [x_next]
            else:
                x0_preds.append(x0_t)
                xs.append(x_next)
        xs = rearrange(torch.stack(xs, dim=0), "t b c h w -> b t h w c")
        x0_preds = rearrange(torch.stack(x0_preds, dim=0), "t b c h w -> b t h w c")
        return xs, x0_preds

    # ---------------------------------------------------------------- evaluation
    def validation_step(self, val_batch, batch_idx):
        raise NotImplementedError("PlDdim.validation_step uses the DDIM sampler configuration; use test_step")

    def test_step(self, test_batch, test_idx):                        # :372-533, the metrics that do not need plots
        h_unnorm, dx, dt, u_unnorm = test_batch
        self.h_ch = h_ch = h_unnorm.shape[-1]
        self.u_ch = u_ch = u_unnorm.shape[-1]
        state_gt = self.data_transform(h_unnorm, u_unnorm)
        h, u = state_gt[..., 0:h_ch], state_gt[..., h_ch:u_ch + h_ch]
        sp = self.test_sparams
        n_samples = sp.n_samples
        state_gt_rep = state_gt.repeat(n_samples, 1, 1, 1)
        if sp.type == "edm":                                             # :398-402
            xs = self.sample_edm(state_gt_rep[..., 0:h_ch], state_gt_rep[..., h_ch:u_ch + h_ch], sp,
                                 return_last=sp.return_last, guide_dx=sp.guide_dx)
        else:
            xs, _ = self.sample_with_repeat(state_gt_rep[..., 0:h_ch], state_gt_rep[..., h_ch:u_ch + h_ch], sp,
                                            return_last=sp.return_last, guide_dx=sp.guide_dx)
        xs_mean = torch.mean(rearrange(xs, "(n b) t h w c -> n b t h w c", n=n_samples), dim=0)
        h_last, u_last = xs_mean[:, -1, :, :, 0:h_ch], xs_mean[:, -1, :, :, h_ch:u_ch + h_ch]
        loss_h, loss_u = self.mae_criterion(h_last, h), self.mae_criterion(u_last, u)
        h_un, u_un = self.inverse_data_transform(h_last, u_last)
        loss_h_un, loss_u_un = self.mae_criterion(h_un, h_unnorm), self.mae_criterion(u_un, u_unnorm)
        hu_un = torch.cat([h_un, u_un], dim=-1)
        mask = torch.ones_like(hu_un)
        if sp.n_time_h > 0:
            mask[:, :sp.n_time_h, :, :h_ch] = 0.0
        if sp.n_time_u > 0:
            mask[:, :sp.n_time_u, :, h_ch:u_ch + h_ch] = 0.0
        loss_hu_un = self.mae_criterion_mask(hu_un, torch.cat([h_unnorm, u_unnorm], dim=-1), mask)
        corr_hu = self.correlation(xs_mean[:, -1], state_gt)
        logs = dict(test_mae_h=loss_h, test_mae_u=loss_u, test_mae_h_un=loss_h_un, test_mae_u_un=loss_u_un,
                    test_mae_hu_un=loss_hu_un, test_corr_h=torch.mean(corr_hu[0:h_ch]),
                    test_corr_u=torch.mean(corr_hu[h_ch:u_ch + h_ch]))
        n_all = h.shape[1]
        if sp.n_time_h < n_all and sp.n_time_h > 0:
            logs["test_h_known"] = self.mae_criterion(h_last[:, :sp.n_time_h], h[:, :sp.n_time_h])
        if n_all > sp.n_time_u > 0:
            logs["test_u_known"] = self.mae_criterion(u_last[:, :sp.n_time_u], u[:, :sp.n_time_u])
        n_batch = len(h_unnorm)
        logs["test_pde_loss"] = self.get_pde_loss(xs[:, -1], clamp_loss=False, do_rearrange=False) / n_samples / n_batch
        logs["test_pde_loss_gt"] = self.get_pde_loss(state_gt, clamp_loss=False, do_rearrange=False) / n_batch
        for k, v in logs.items():
            self.log(k, v, prog_bar=True, on_epoch=True, on_step=False, sync_dist=True)
        traj = rearrange(xs[:, -1], "(n b) h w c -> b h w n c", n=n_samples).unsqueeze(dim=1)
        return {"loss_h": loss_h, "loss_u": loss_u, "loss_hu_un": loss_hu_un, "traj": traj, "gt": state_gt}
