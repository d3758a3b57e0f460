"""Host mirror of the reference's PDE-residual modules on the K6 kernels (csrc/pde.cu):

    SweFvLoss, DarcyLoss        models/pde_loss.py:19-86, :89-246
    get_pde_loss_function       models/loss_helper.py:14-41

Same class names, constructor arguments and `forward(pred, gt, normalizer_h, normalizer_u, return_d, calc_prob,
clamp_loss)` contract (pred / gt: un-normalised `[B, T, X, 2]`).  The module-level callers (`get_pde_loss`,
`get_dx_pde` in mcedm.py / cond_edm.py) use `residual_sum` / `gradient` instead, which take the two *normalised*
channel planes in whatever layout and dtype the sampler holds them (float64 NCHW state, float32 condition) and fold the
cast, the inverse normalisation and the reduction into the one launch.

There is no CPU path: CPU tensors raise.  Not implemented (raise NotImplementedError): `flip_xy`, the Darcy gradient
(`DarcyLoss(return_d=True)`), `unroll_loss`.
"""
from __future__ import annotations

import sys

import torch
from torch import nn

from . import _lib as L


def _plane(t):
    """(tensor, f64 flag, strides) of a [B,T,X] view (any strides, float32 / float64)."""
    if t.dim() != 3:
        raise ValueError(f"expected a [B,T,X] plane, got {tuple(t.shape)}")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    if not t.is_cuda:
        raise L.McedmError("the PDE residual kernels need CUDA tensors: there is no CPU fallback")
    return t, int(t.dtype == torch.float64), L.ll_array(t.stride())


def _stat(v):
    v = torch.as_tensor(v)
    if v.numel() != 1:
        raise NotImplementedError("per-channel normalisation statistics with more than one entry")
    return float(v.reshape(-1)[0])


_STAT_CACHE = {}


def _norm_args(normalizer_h, normalizer_u):
    """(h_div, h_sub, u_div, u_sub) as python floats.  Reading a CUDA scalar synchronises with the device, so the values
    are cached per (buffer objects, their versions); the cache holds the buffers themselves, so an id can never be
    re-used by another tensor while its entry is alive, and in-place updates (version bump) or `set_stats` (new
    objects) miss."""
    ts = (normalizer_h.divide, normalizer_h.subtract, normalizer_u.divide, normalizer_u.subtract)
    key = tuple((id(t), t._version) for t in ts)
    hit = _STAT_CACHE.get(key)
    if hit is None or any(a is not b for a, b in zip(hit[0], ts)):
        if len(_STAT_CACHE) > 64:
            _STAT_CACHE.clear()
        hit = _STAT_CACHE[key] = (ts, tuple(_stat(t) for t in ts))
    return hit[1]


class SweFvLoss(nn.Module):
    """PDE loss function for shallow water equations using the finite volume (FORCE) method."""

    def __init__(self, Tn=0.128, x_min=-2.5, x_max=2.5, n_ghosts=2, reduction="none", flip_xy=False):
        super().__init__()
        if n_ghosts != 2:
            raise NotImplementedError("the FORCE kernels are written for n_ghosts = 2")
        self.flip_xy = flip_xy
        self.g = 1.0
        self.Tn = Tn
        self.x_min = x_min
        self.x_max = x_max
        self.n_ghosts = n_ghosts
        self.eps = 1e-8

    def gen_x(self, nx, s_t=None):                                   # pde_loss.py:104-118 (host, float32 linspace)
        step = (self.x_max - self.x_min) / nx
        n_ghosts = self.n_ghosts
        nx += 2 * n_ghosts
        if nx % 2 == 0:
            x = torch.linspace(self.x_min + step / 2 - step * n_ghosts, self.x_max - step / 2 + step * n_ghosts, nx)
        else:
            x = torch.linspace(self.x_min - step * n_ghosts, self.x_max + step * n_ghosts, nx)
        return x

    def _grid(self, T, X):
        """(float32(0.5*dt) as a python float, dx): dt = Tn / n_times (:207), dx = x[1] - x[0] (:136-137)."""
        key = (T, X)
        cache = self.__dict__.setdefault("_grid_cache", {})
        if key not in cache:
            x = self.gen_x(X)
            cache[key] = (0.5 * (self.Tn / T), float(x[1] - x[0]))
        return cache[key]

    def _check(self):
        if self.flip_xy:
            raise NotImplementedError("flip_xy is not supported by the PDE residual kernels")

    # ---- fused entry points on normalised planes -----------------------------------------------
    def residual(self, h, u, normalizer_h, normalizer_u, gt=None, apply_norm=True, want_matrix=False):
        """h, u: [B,T,X] planes.  Returns (loss matrix [B,T,X,2] float32 or None, 0-dim float64 sum)."""
        self._check()
        h, hf, hs = _plane(h)
        u, uf, us = _plane(u)
        B, T, X = h.shape
        half_dt, dx = self._grid(T, X)
        hd, hsub, ud, usub = _norm_args(normalizer_h, normalizer_u)
        dev = h.device
        loss = torch.empty(B, T, X, 2, device=dev, dtype=torch.float32) if want_matrix else None
        rows = torch.empty(B * T, device=dev, dtype=torch.float64)
        total = torch.empty((), device=dev, dtype=torch.float64)
        if gt is not None:
            gt = gt.to(torch.float32).contiguous()
        L.check(L.lib().mcedm_swe_fv_loss(L.ptr(h), hf, hs, L.ptr(u), uf, us, int(apply_norm), hd, hsub, ud, usub,
                                          L.ptr(gt), B, T, X, half_dt, dx, self.g, L.ptr(loss), L.ptr(rows),
                                          L.ptr(total), L.stream_ptr()), "swe_fv_loss")
        L.LAUNCHES[0] += 1                                            # the fixed-order sum is a second launch
        return loss, total

    def gradient(self, h, u, normalizer_h, normalizer_u, gt=None, apply_norm=True, mode=0):
        """d mean(residual matrix) / d (un-normalised h, u).  mode 0: [B,T,X,2]; 1: channel mean [B,T,X]; 2: sum."""
        self._check()
        h, hf, hs = _plane(h)
        u, uf, us = _plane(u)
        B, T, X = h.shape
        half_dt, dx = self._grid(T, X)
        hd, hsub, ud, usub = _norm_args(normalizer_h, normalizer_u)
        out = torch.empty((B, T, X, 2) if mode == 0 else (B, T, X), device=h.device, dtype=torch.float32)
        if gt is not None:
            gt = gt.to(torch.float32).contiguous()
        L.check(L.lib().mcedm_swe_fv_grad(L.ptr(h), hf, hs, L.ptr(u), uf, us, int(apply_norm), hd, hsub, ud, usub,
                                          L.ptr(gt), B, T, X, half_dt, dx, self.g, mode, L.ptr(out), L.stream_ptr()),
                "swe_fv_grad")
        return out

    # ---- the reference's module contract (pde_loss.py:227-246) ---------------------------------
    def calculate_loss(self, pred, gt, normalizer_h, normalizer_u):
        return self.residual(pred[..., 0], pred[..., 1], normalizer_h, normalizer_u, gt=gt, apply_norm=False,
                             want_matrix=True)[0]

    def forward(self, pred, gt, normalizer_h, normalizer_u, return_d=False, calc_prob=False, clamp_loss=False):
        if return_d:
            return self.gradient(pred[..., 0], pred[..., 1], normalizer_h, normalizer_u, gt=gt, apply_norm=False)
        loss = self.calculate_loss(pred, gt, normalizer_h, normalizer_u)
        if clamp_loss:
            loss = torch.clamp(loss, max=1.0)
        return loss

    def unroll_loss(self, *a, **k):
        raise NotImplementedError("unroll_loss (simulator roll-out metric) is outside the hot path")


class SweSimulatorLoss(SweFvLoss):
    """loss_helper.py:5-11: the FV loss stands in when the PyClaw-based simulator loss is unavailable."""


class DarcyLoss(nn.Module):
    """PDE loss function for the Darcy flow equation (central differences, beta = 1)."""

    def __init__(self, reduction="none", flip_xy=False):
        super().__init__()
        self.flip_xy = flip_xy
        self.D = 1.0
        self.eps = 1e-8

    def residual(self, a, u, normalizer_h, normalizer_u, apply_norm=True, want_matrix=False):
        """a, u: [B,S,S] planes (permeability, solution).  Returns (loss [B,S-4,S-4] / ((S-4)^2) or None, sum)."""
        if self.flip_xy:
            a, u = u, a
            normalizer_h, normalizer_u = normalizer_u, normalizer_h
        a, af, as_ = _plane(a)
        u, uf, us = _plane(u)
        B, S, S2 = a.shape
        if S != S2:
            raise ValueError("DarcyLoss expects square fields")
        ad, asub, ud, usub = _norm_args(normalizer_h, normalizer_u)
        n = S - 4
        loss = torch.empty(B, n, n, device=a.device, dtype=torch.float32) if want_matrix else None
        rows = torch.empty(B * n, device=a.device, dtype=torch.float64)
        total = torch.empty((), device=a.device, dtype=torch.float64)
        L.check(L.lib().mcedm_darcy_loss(L.ptr(a), af, as_, L.ptr(u), uf, us, int(apply_norm), ad, asub, ud, usub, B, S,
                                         self.D, L.ptr(loss), L.ptr(rows), L.ptr(total), L.stream_ptr()), "darcy_loss")
        L.LAUNCHES[0] += 1
        return loss, total

    def gradient(self, *a, **k):
        raise NotImplementedError("the Darcy residual gradient (guidance) has no sm_100a kernel")

    def forward(self, pred, gt, normalizer_h, normalizer_u, return_d=False, calc_prob=False, clamp_loss=False):
        if return_d:
            return self.gradient()
        # flip_xy: residual() swaps the two planes, which is flip_state (pde_loss.py:6-16) for a 2-channel field
        loss = self.residual(pred[..., 0], pred[..., 1], normalizer_h, normalizer_u, apply_norm=False,
                             want_matrix=True)[0]
        if clamp_loss:
            loss = torch.clamp(loss, max=1.0)
        return loss


def get_pde_loss_function(system, flip_xy, Tn_mult=1.0):
    """loss_helper.py:14-41 (the undefined `ReactorLoss` branch raises NameError there; here NotImplementedError)."""
    # the reference prints this line to stdout (loss_helper.py:15); stderr here: bench.py's stdout is one JSON line
    print(f"PDE error: system = {system}", file=sys.stderr)
    if system == "swe_per":
        Tn = 0.128 * Tn_mult
        return (SweFvLoss(Tn=Tn, x_min=-0.5, x_max=0.5, flip_xy=flip_xy),
                SweSimulatorLoss(Tn=Tn, x_min=-0.5, x_max=0.5, flip_xy=flip_xy))
    if system == "darcy":
        return DarcyLoss(flip_xy=flip_xy), DarcyLoss(flip_xy=flip_xy)
    if system == "reactor":
        raise NotImplementedError("ReactorLoss is undefined in the reference (loss_helper.py:30-31)")
    Tn = 1.28 * Tn_mult
    return SweFvLoss(Tn=Tn, flip_xy=flip_xy), SweSimulatorLoss(Tn=Tn, flip_xy=flip_xy)
