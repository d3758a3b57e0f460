"""CPU oracle of the PDE-residual path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module.

Restates, in numpy float32 arithmetic (one IEEE-754 rounding per operation, in the order in which the
reference's torch expressions evaluate), the reference functions

    SweFvLoss.f_t_swp1d / set_boundary / gen_x      models/pde_loss.py:104-165   (FORCE finite-volume step)
    SweFvLoss.calculate_loss / get_scaling / forward models/pde_loss.py:199-246  (residual matrix, d/dpred)
    DarcyLoss.calculate_loss / forward               models/pde_loss.py:30-86
    PlMcedm.get_pde_loss                             models/mcedm.py:468-499      (joint (h,u) sample)
    PlCondDdim.get_pde_loss / get_dx_pde             models/ddim.py:1388-1450     (h = condition, u = sample)
    get_pde_loss_function                            models/loss_helper.py:14-41
    Normalizer(inverse=True)                         models/normalizer.py:25-27

The gradient (`return_d=True`, models/pde_loss.py:231-242) is autograd in the reference; here it is the analytic
adjoint of the FORCE step evaluated in float64 (tolerance-checked, not bit-checked).

Pinning: the reference ships no tests for this path; `tests/golden/pde.pt` holds outputs of the UNMODIFIED
reference (tests/golden/make_golden_pde.py) and tests/test_oracle_golden.py checks every function here against it
(residual matrices bit-exact, gradients to 1e-5 of their max).
"""
from __future__ import annotations

import numpy as np
import torch

f32 = np.float32


def pde_params(system: str, n_x: int, n_t: int, Tn_mult: float = 1.0):
    """(Tn, x_min, x_max, dt, dx) of get_pde_loss_function (loss_helper.py:14-41) + gen_x (pde_loss.py:104-118).
    dx = x[1] - x[0] of the float32 `torch.linspace` grid with ghost cells, as f_t_swp1d reads it (:136-137)."""
    if system == "swe_per":
        Tn, x_min, x_max = 0.128 * Tn_mult, -0.5, 0.5
    else:  # "swe" and the default branch
        Tn, x_min, x_max = 1.28 * Tn_mult, -2.5, 2.5
    step = (x_max - x_min) / n_x
    n_g = 2
    nx = n_x + 2 * n_g
    if nx % 2 == 0:
        x = torch.linspace(x_min + step / 2 - step * n_g, x_max - step / 2 + step * n_g, nx)
    else:
        x = torch.linspace(x_min - step * n_g, x_max + step * n_g, nx)
    dx = float(x[1] - x[0])
    dt = Tn / n_t
    return Tn, x_min, x_max, dt, dx


def unnormalize(x, divide, subtract):
    """Normalizer.forward(inverse=True): x * divide + subtract in float32 (normalizer.py:26-27)."""
    return (np.asarray(x, dtype=f32) * f32(divide)).astype(f32) + f32(subtract)


def swe_fv_step(h_in, u_in, dt, dx, g=1.0, eps=1e-8):
    """One FORCE step of every row; h_in, u_in float32 [..., nx] -> (h_next, u_next) [..., nx]  (pde_loss.py:129-165)."""
    h = np.pad(np.asarray(h_in, f32), [(0, 0)] * (h_in.ndim - 1) + [(2, 2)], mode="edge")
    u = np.pad(np.asarray(u_in, f32), [(0, 0)] * (u_in.ndim - 1) + [(2, 2)], mode="edge")
    q = u * h                                                   # hu = s[...,1] * s[...,0]
    c = f32(0.5 * dt)                                           # python float 0.5*dt, cast to the tensor dtype
    dx = f32(dx)
    e = f32(eps)
    hg = f32(0.5 * g)

    def flux(qq, hh):                                           # hu**2 / (h + eps) + 0.5*g*h**2
        return (qq * qq) / (hh + e) + hg * (hh * hh)

    def half(a, b):                                             # 0.5*(a[:-1]+a[1:]) - 0.5*dt*(b[1:]-b[:-1])/dx
        return f32(0.5) * (a[..., :-1] + a[..., 1:]) - (c * (b[..., 1:] - b[..., :-1])) / dx

    F = flux(q, h)
    hm = half(h, q)
    qm = half(q, F)
    hn = half(hm, qm)
    G = flux(qm, hm)
    qn = half(qm, G)
    H = hn[..., 1:-1]
    U = qn[..., 1:-1] / (H + e)
    return H, U


def swe_fv_loss_matrix(pred, gt, scale_h, scale_u, dt, dx):
    """SweFvLoss.calculate_loss (pde_loss.py:199-215): pred, gt float32 [B,T,X,2] un-normalised -> [B,T,X,2]."""
    pred = np.asarray(pred, f32)
    gt = np.asarray(gt, f32)
    H, U = swe_fv_step(pred[..., 0], pred[..., 1], dt, dx)
    nxt = np.stack([H, U], axis=-1)
    with_ic = np.concatenate([pred[:, 0:1], nxt[:, :-1]], axis=1)
    with_ic = np.where(np.isnan(with_ic), f32(0), with_ic)
    scale = np.array([f32(scale_h) * f32(scale_h), f32(scale_u) * f32(scale_u)], dtype=f32)
    d = with_ic - gt
    return (d * d) / scale


def swe_fv_grad(pred, gt, scale_h, scale_u, dt, dx, g=1.0, eps=1e-8):
    """d mean(loss_matrix) / d pred with gt held constant (SweFvLoss.forward, return_d=True, pde_loss.py:231-242);
    analytic adjoint in float64 of the float32 forward values, NaN entries set to 0.  [B,T,X,2] float32."""
    p = np.asarray(pred, f32).astype(np.float64)
    gtd = np.asarray(gt, f32).astype(np.float64)
    B, T, X, _ = p.shape
    n_el = p.size
    sc = np.array([float(f32(scale_h)) ** 2, float(f32(scale_u)) ** 2])
    cdx = 0.5 * dt / float(f32(dx))
    h = np.pad(p[..., 0], [(0, 0), (0, 0), (2, 2)], mode="edge")
    u = np.pad(p[..., 1], [(0, 0), (0, 0), (2, 2)], mode="edge")
    q = u * h
    F = q * q / (h + eps) + 0.5 * g * h * h
    hm = 0.5 * (h[..., :-1] + h[..., 1:]) - cdx * (q[..., 1:] - q[..., :-1])
    qm = 0.5 * (q[..., :-1] + q[..., 1:]) - cdx * (F[..., 1:] - F[..., :-1])
    G = qm * qm / (hm + eps) + 0.5 * g * hm * hm
    hn = 0.5 * (hm[..., :-1] + hm[..., 1:]) - cdx * (qm[..., 1:] - qm[..., :-1])
    qn = 0.5 * (qm[..., :-1] + qm[..., 1:]) - cdx * (G[..., 1:] - G[..., :-1])
    H = hn[..., 1:-1]
    Q = qn[..., 1:-1]
    U = Q / (H + eps)
    # residual of row t+1 is fed by the step of row t; the last row's step is dropped, row 0 is compared directly
    gH = np.zeros_like(H)
    gU = np.zeros_like(U)
    rH = H[:, :-1] - gtd[:, 1:, :, 0]
    rU = U[:, :-1] - gtd[:, 1:, :, 1]
    gH[:, :-1] = np.where(np.isnan(H[:, :-1]), 0.0, 2.0 * rH / sc[0] / n_el)
    gU[:, :-1] = np.where(np.isnan(U[:, :-1]), 0.0, 2.0 * rU / sc[1] / n_el)
    # U = Q/(H+eps)
    g_qn = np.zeros_like(qn)
    g_hn = np.zeros_like(hn)
    g_qn[..., 1:-1] = gU / (H + eps)
    g_hn[..., 1:-1] = gH - gU * Q / (H + eps) ** 2

    def half_adj(g_out, n_in):
        """adjoint of out = 0.5*(a[:-1]+a[1:]) - cdx*(b[1:]-b[:-1]) -> (g_a, g_b) of length n_in"""
        g_a = np.zeros(g_out.shape[:-1] + (n_in,))
        g_b = np.zeros_like(g_a)
        g_a[..., :-1] += 0.5 * g_out
        g_a[..., 1:] += 0.5 * g_out
        g_b[..., 1:] -= cdx * g_out
        g_b[..., :-1] += cdx * g_out
        return g_a, g_b

    g_hm, g_qm = half_adj(g_hn, hm.shape[-1])
    a, g_G = half_adj(g_qn, hm.shape[-1])
    g_qm += a
    g_qm += g_G * 2.0 * qm / (hm + eps)
    g_hm += g_G * (-(qm * qm) / (hm + eps) ** 2 + g * hm)
    g_h, g_q = half_adj(g_hm, h.shape[-1])
    a, g_F = half_adj(g_qm, h.shape[-1])
    g_q += a
    g_q += g_F * 2.0 * q / (h + eps)
    g_h += g_F * (-(q * q) / (h + eps) ** 2 + g * h)
    g_u = g_q * h
    g_h += g_q * u

    def fold(gp):                                               # adjoint of the replicate padding
        out = gp[..., 2:-2].copy()
        out[..., 0] += gp[..., 0] + gp[..., 1]
        out[..., -1] += gp[..., -1] + gp[..., -2]
        return out

    grad = np.stack([fold(g_h), fold(g_u)], axis=-1)
    grad[:, 0] += 2.0 * (p[:, 0] - gtd[:, 0]) / sc / n_el       # row 0 is compared with gt directly
    grad = np.where(np.isnan(grad), 0.0, grad)
    return grad.astype(f32)


def darcy_loss_matrix(pred, D=1.0):
    """DarcyLoss.calculate_loss + the /(t*n) of forward (pde_loss.py:30-56, :81-84): pred float32 [B,S,S,2] (a, u)
    -> [B,S-4,S-4]."""
    pred = np.asarray(pred, f32)
    size = pred.shape[1]
    a = pred[..., 0]
    u = pred[..., 1]
    dx = D / size                                               # python float
    two_dx = f32(2 * dx)
    ux = (u[:, 2:, 1:-1] - u[:, :-2, 1:-1]) / two_dx
    uy = (u[:, 1:-1, 2:] - u[:, 1:-1, :-2]) / two_dx
    a = a[:, 1:-1, 1:-1]
    aux = a * ux
    auy = a * uy
    auxx = (aux[:, 2:, 1:-1] - aux[:, :-2, 1:-1]) / two_dx
    auyy = (auy[:, 1:-1, 2:] - auy[:, 1:-1, :-2]) / two_dx
    Du = -(auxx + auyy)
    r = Du - f32(1.0)
    loss = r * r
    t, n = loss.shape[1:]
    return loss / f32(t * n)


# ------------------------------------------------------------------------------------------------
# module-level wrappers
# ------------------------------------------------------------------------------------------------
def get_pde_loss(h_norm, u_norm, stats, system, flip_xy=False):
    """PlMcedm.get_pde_loss / PlCondDdim.get_pde_loss with clamp_loss=False, x_gt_unnorm=None, reduce=True
    (mcedm.py:468-499, ddim.py:1388-1422): h_norm, u_norm [B,T,X] normalised (any float dtype; cast to float32 first).
    Returns (loss matrix, float64 sum of its entries)."""
    h = unnormalize(np.asarray(h_norm).astype(f32), stats["input_std"], stats["input_mean"])
    u = unnormalize(np.asarray(u_norm).astype(f32), stats["target_std"], stats["target_mean"])
    if system == "darcy":
        m = darcy_loss_matrix(np.stack([u, h] if flip_xy else [h, u], axis=-1))
    else:
        if flip_xy:
            raise NotImplementedError("flip_xy for the SWE residual")
        _, _, _, dt, dx = pde_params(system, h.shape[2], h.shape[1])
        x = np.stack([h, u], axis=-1)
        m = swe_fv_loss_matrix(x, x, stats["input_std"], stats["target_std"], dt, dx)
    return m, float(m.astype(np.float64).sum())


def get_dx_pde_cond(h_norm, u_norm, stats, system, calc_prob=True):
    """PlCondDdim.get_dx_pde (ddim.py:1424-1450) for the SWE systems: gradient of the mean residual w.r.t. the
    un-normalised (h, u), then mean over the two channels (calc_prob=True, keepdim) or their sum.  -> [B,1|-,T,X]"""
    h = unnormalize(np.asarray(h_norm).astype(f32), stats["input_std"], stats["input_mean"])
    u = unnormalize(np.asarray(u_norm).astype(f32), stats["target_std"], stats["target_mean"])
    _, _, _, dt, dx = pde_params(system, h.shape[2], h.shape[1])
    x = np.stack([h, u], axis=-1)
    g = swe_fv_grad(x, x, stats["input_std"], stats["target_std"], dt, dx)
    if calc_prob:
        return ((g[..., 0] + g[..., 1]) / f32(2.0))[:, None]
    return g[..., 0] + g[..., 1]
