"""Synthetic data with the reference's batch contract, and its mask generators.

The reference reads PyClaw-generated HDF5 files (datamodules/h5_dataset.py, pl_datamodule.py); neither
h5py nor the data exist here, and BASELINE.json quotes every metric on *synthetic* fields of the same
shape.  This module provides:

  * field generators with the shape and statistics of the reference data sets
      swe_periodic_fields  — 128 x 128 (t, x) fields of (h, u), periodic 1-D shallow water; initial
                             condition exactly as generate/src/sim_dam_break_1d.py:476-490
                             (7-mode random Fourier series rescaled to h0 in [1, 2], seed = item index),
                             evolved with the linearised SWE (two travelling waves, g = 1) instead of PyClaw;
      dam_break_fields     — Gaussian perturbation on a constant state, parameters as
                             generate/gen_dam_break_1d.py:66-73, same linear evolution with advection;
      darcy_fields         — piecewise-constant permeability a in {0.1, 1.0} and a smooth u >= 0
                             (PDEBench Darcy shape, preprocess_darcy.py);
  * the mask generators of datamodules/h5_dataset.py (:232-255 HDF5MaskDataset.sample_mask,
    :306-393 HDF5TimeMaskDataset) with identical RNG consumption (`torch.rand(1)` / `torch.randint`
    per item), mask == 1 meaning "missing, to be generated";
  * datamodules exposing the attributes PlMcedm reads (`get_norm_stats`, `down_factor`, `down_interp`)
    and yielding the reference batch tuples:
        mask datamodules : (h[B,128,128,1], t_grid, x_grid, u[B,128,128,1], mask[B,128,128,2] | {'u','h'[,'hu']})
        plain datamodule : (h, dx, dt, u)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset, DistributedSampler

from .config import AttrDict


# ---------------------------------------------------------------------------------------------
# field generators
# ---------------------------------------------------------------------------------------------
def swe_periodic_fields(n: int, res: int = 128, first_seed: int = 0, g: float = 1.0,
                        t_end: float = 0.128) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (h, u), each float32 [n, res_t, res_x, 1]."""
    x = -0.5 + (np.arange(res) + 0.5) / res
    t = np.arange(res) * (t_end / res)
    ks = np.arange(-3, 4)
    hs, us = [], []
    for i in range(n):
        rng = np.random.RandomState(first_seed + i)
        lam, gam = rng.randn(7), rng.randn(7)

        def h_hat(xx):
            ph = 2 * np.pi * ks[None, :] * xx[..., None]
            return (lam * np.cos(ph) + gam * np.sin(ph)).sum(-1)

        base = h_hat(x)
        lo, hi = base.min(), base.max()
        h0 = lambda xx: 1.0 + (h_hat(xx) - lo) / (hi - lo)  # noqa: E731
        hbar = h0(x).mean()
        c = np.sqrt(g * hbar)
        xr = x[None, :] - c * t[:, None]
        xl = x[None, :] + c * t[:, None]
        f_r, f_l = 0.5 * (h0(xr) - hbar), 0.5 * (h0(xl) - hbar)
        hs.append(hbar + f_r + f_l)
        us.append((c / hbar) * (f_r - f_l))
    h = np.stack(hs).astype(np.float32)[..., None]
    u = np.stack(us).astype(np.float32)[..., None]
    return h, u


def dam_break_fields(n: int, res: int = 128, first_seed: int = 0, g: float = 1.0,
                     t_end: float = 1.28) -> Tuple[np.ndarray, np.ndarray]:
    x = -2.5 + 5.0 * (np.arange(res) + 0.5) / res
    t = np.arange(res) * (t_end / res)
    hs, us = [], []
    for i in range(n):
        rng = np.random.RandomState(first_seed + i)
        H, eps = rng.uniform(1.2, 5.2), rng.uniform(0.05, 1.0)
        x0, s, u0 = rng.uniform(-1, 1), rng.uniform(0.2, 2.0), rng.uniform(-2.2, 2.2)
        c = np.sqrt(g * H)
        bump = lambda xx: 0.5 * eps * np.exp(-0.5 * (xx - x0) ** 2 / s ** 2)  # noqa: E731
        f_r = bump(x[None, :] - (u0 + c) * t[:, None])
        f_l = bump(x[None, :] - (u0 - c) * t[:, None])
        hs.append(H + f_r + f_l)
        us.append(u0 + (c / H) * (f_r - f_l))
    return np.stack(hs).astype(np.float32)[..., None], np.stack(us).astype(np.float32)[..., None]


def darcy_fields(n: int, res: int = 128, first_seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    yy, xx = np.meshgrid((np.arange(res) + 0.5) / res, (np.arange(res) + 0.5) / res, indexing="ij")
    k = np.fft.fftfreq(res)[:, None] ** 2 + np.fft.fftfreq(res)[None, :] ** 2
    lowpass = np.exp(-k * (res / 6.0) ** 2)
    a_s, u_s = [], []
    for i in range(n):
        rng = np.random.RandomState(first_seed + i)
        field = np.fft.ifft2(np.fft.fft2(rng.randn(res, res)) * lowpass).real
        a = np.where(field > 0, 1.0, 0.1)
        src = np.fft.ifft2(np.fft.fft2(1.0 / a) * lowpass).real
        u = np.sin(np.pi * xx) * np.sin(np.pi * yy) * src * 0.05
        a_s.append(a)
        u_s.append(np.maximum(u, 0.0))
    return np.stack(a_s).astype(np.float32)[..., None], np.stack(u_s).astype(np.float32)[..., None]


_FIELDS = {"swe": dam_break_fields, "swe_per": swe_periodic_fields, "darcy": darcy_fields}


# ---------------------------------------------------------------------------------------------
# mask generators (mask == 1 -> missing)
# ---------------------------------------------------------------------------------------------
def sample_mask(inp: torch.Tensor, target: torch.Tensor, is_train: bool):
    """HDF5MaskDataset.sample_mask: train = one coin flip per item choosing which channel is missing;
    eval = both variants, keyed by the name of the missing variable."""
    zi, zt, oi, ot = torch.zeros_like(inp), torch.zeros_like(target), torch.ones_like(inp), torch.ones_like(target)
    if is_train:
        if torch.rand(1) > 0.5:
            return torch.cat([zi, ot], dim=-1)
        return torch.cat([oi, zt], dim=-1)
    return {"u": torch.cat([zi, ot], dim=-1), "h": torch.cat([oi, zt], dim=-1)}


def sample_time_mask(inp: torch.Tensor, target: torch.Tensor, is_train: bool, add_time_masks: bool = False):
    """HDF5TimeMaskDataset: train = variable choice (0.4 / 0.4 / 0.2) OR-ed with a random observation
    horizon per channel, t_max in [res/2, res]; eval = {'u','h'} or, with add_time_masks, the three
    half-horizon masks {'hu','u','h'}."""
    ic = inp.shape[-1]
    if is_train:
        var = torch.rand(1)
        zi = torch.zeros_like(inp, dtype=torch.bool)
        zt = torch.zeros_like(target, dtype=torch.bool)
        if var <= 0.4:
            mask_var = torch.cat([zi, ~zt], dim=-1)
        elif var <= 0.8:
            mask_var = torch.cat([~zi, zt], dim=-1)
        else:
            mask_var = torch.cat([zi, zt], dim=-1)
        res = inp.shape[0]
        t1 = res // 2 + torch.randint(res // 2 + 1, (1,))
        t2 = res // 2 + torch.randint(res // 2 + 1, (1,))
        horizon = torch.ones_like(mask_var, dtype=torch.bool)
        horizon[:t1, :, :ic] = False
        horizon[:t2, :, ic:] = False
        return (mask_var | horizon).float()
    masks = sample_mask(inp, target, False)
    if add_time_masks:
        half = int(0.5 * inp.shape[0])
        late_i, late_t = torch.zeros_like(inp), torch.zeros_like(target)
        late_i[half:] = 1
        late_t[half:] = 1
        masks = {"hu": torch.cat([late_i, late_t], dim=-1),
                 "u": torch.cat([late_i, torch.ones_like(target)], dim=-1),
                 "h": torch.cat([torch.ones_like(inp), late_t], dim=-1)}
    return masks


# ---- the same masks as two integers per item (device-side generation) -------------------------------
# Every training mask above is "channel c is missing from time row r_c on": r = 0 (whole channel missing), r = res
# (observed), or the random horizon t_max.  `*_rows` make EXACTLY the RNG calls of their tensor-valued counterparts (same
# generator state afterwards) and return (r_h, r_u); `expand_mask_rows` is the host restatement of what
# mcedm_mcedm_prep_rows expands on the device (bit-identical to sample_mask / sample_time_mask for h_ch = u_ch = 1).
def sample_mask_rows(res: int) -> torch.Tensor:
    """HDF5MaskDataset.sample_mask(is_train=True) (h5_dataset.py:236-243) as observation rows."""
    if torch.rand(1) > 0.5:
        return torch.tensor([res, 0], dtype=torch.int32)              # u missing
    return torch.tensor([0, res], dtype=torch.int32)                  # h missing


def sample_time_mask_rows(res: int) -> torch.Tensor:
    """HDF5TimeMaskDataset training mask (h5_dataset.py:326-345) as observation rows."""
    var = torch.rand(1)
    t1 = int(res // 2 + torch.randint(res // 2 + 1, (1,)))
    t2 = int(res // 2 + torch.randint(res // 2 + 1, (1,)))
    if var <= 0.4:
        return torch.tensor([t1, 0], dtype=torch.int32)               # u missing entirely, h observed up to its horizon
    if var <= 0.8:
        return torch.tensor([0, t2], dtype=torch.int32)
    return torch.tensor([t1, t2], dtype=torch.int32)


def expand_mask_rows(rows: torch.Tensor, res_t: int, res_x: int) -> torch.Tensor:
    """[.., 2] observation rows -> fp32 mask [.., res_t, res_x, 2] (1 = missing)."""
    t = torch.arange(res_t, device=rows.device).reshape(res_t, 1, 1)
    return (t >= rows.reshape(*rows.shape[:-1], 1, 1, 2)).float().expand(*rows.shape[:-1], res_t, res_x, 2).contiguous()


# ---------------------------------------------------------------------------------------------
# datasets / datamodules
# ---------------------------------------------------------------------------------------------
class _FieldDataset(Dataset):
    def __init__(self, h: np.ndarray, u: np.ndarray, mask_mode: str, is_train: bool, add_time_masks: bool,
                 return_grid: bool):
        self.h, self.u = torch.from_numpy(h), torch.from_numpy(u)
        self.mask_mode, self.is_train, self.add_time_masks = mask_mode, is_train, add_time_masks
        res_t, res_x = h.shape[1], h.shape[2]
        if return_grid:
            tt = torch.linspace(0, 1, res_t).reshape(res_t, 1, 1).expand(res_t, res_x, 1)
            xx = torch.linspace(0, 1, res_x).reshape(1, res_x, 1).expand(res_t, res_x, 1)
            self.g0, self.g1 = tt.contiguous(), xx.contiguous()
        else:
            self.g0, self.g1 = torch.tensor(1.0 / res_x), torch.tensor(1.0 / res_t)

    def __len__(self):
        return self.h.shape[0]

    def __getitem__(self, idx):
        inp, tar = self.h[idx], self.u[idx]
        if self.mask_mode == "none":
            return inp, self.g0, self.g1, tar
        if self.is_train and getattr(self, "device_masks", False) and inp.shape[-1] == 1 and tar.shape[-1] == 1:
            # training masks as observation rows: the mask tensor is expanded on the device (mcedm_mcedm_prep_rows)
            rows = sample_time_mask_rows(inp.shape[0]) if self.mask_mode == "time" else sample_mask_rows(inp.shape[0])
            return inp, self.g0, self.g1, tar, rows
        if self.mask_mode == "time":
            mask = sample_time_mask(inp, tar, self.is_train, self.add_time_masks)
        else:
            mask = sample_mask(inp, tar, self.is_train)
        return inp, self.g0, self.g1, tar, mask


class SyntheticDatamodule:
    """Stand-in for HDF5Datamodule (pl_datamodule.py:10-218): batch = (h, dx, dt, u)."""

    mask_mode = "none"

    def __init__(self, system: str = "swe_per", n_train: int = 64, n_test: int = 16, batch_size: int = 32,
                 test_batch_size: int = 0, num_workers: int = 0, down_factor: int = 1, return_grid: bool = False,
                 add_time_masks: bool = False, resolution: int = 128, **_ignored):
        self.system = system if system in _FIELDS else "swe"
        self.n_train, self.n_test = n_train, n_test
        self.batch_size = batch_size
        self.test_batch_size = test_batch_size if test_batch_size and test_batch_size > 0 else batch_size
        self.down_factor, self.down_interp = down_factor, True
        self.return_grid, self.add_time_masks, self.resolution = return_grid, add_time_masks, resolution
        self.device_masks = bool(_ignored.get("device_masks", False))   # training masks as [B,2] observation rows
        self.eps = 1e-8
        self._built = False

    def _build(self):
        if self._built:
            return
        gen = _FIELDS[self.system]
        self._train = gen(self.n_train, self.resolution, first_seed=0)
        self._test = gen(self.n_test, self.resolution, first_seed=100000)
        h, u = self._train
        t = lambda v: torch.tensor([float(v)], dtype=torch.float32)  # noqa: E731  stats are shape-(1,) arrays
        self.input_mean, self.input_std = t(h.mean()), t(h.std()) + self.eps
        self.target_mean, self.target_std = t(u.mean()), t(u.std()) + self.eps
        self.input_min, self.input_min_max = t(h.min()), t(h.max() - h.min()) + self.eps
        self.target_min, self.target_min_max = t(u.min()), t(u.max() - u.min()) + self.eps
        self._built = True

    def setup(self, stage=None):
        self._build()
        mk = lambda d, tr: _FieldDataset(d[0], d[1], self.mask_mode, tr, self.add_time_masks, self.return_grid)  # noqa
        self.train_dataset, self.val_dataset, self.test_dataset = mk(self._train, True), mk(self._test, False), \
            mk(self._test, False)
        self.train_dataset.device_masks = self.device_masks

    def get_norm_stats(self) -> AttrDict:
        self._build()
        return AttrDict(norm_input=False, norm_target=False, input_mean=self.input_mean, input_std=self.input_std,
                        input_min=self.input_min, input_min_max=self.input_min_max, target_mean=self.target_mean,
                        target_std=self.target_std, target_min=self.target_min, target_min_max=self.target_min_max)

    def _loader(self, ds, bs, shuffle):
        sampler = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            sampler = DistributedSampler(ds, shuffle=shuffle)
            shuffle = False
        return DataLoader(ds, batch_size=bs, shuffle=shuffle, sampler=sampler, num_workers=0, pin_memory=True)

    def train_dataloader(self):
        return self._loader(self.train_dataset, self.batch_size, True)

    def val_dataloader(self):
        return self._loader(self.val_dataset, self.batch_size, False)

    def test_dataloader(self):
        return self._loader(self.test_dataset, self.test_batch_size, False)


class SyntheticMaskDatamodule(SyntheticDatamodule):
    """Stand-in for HDF5MaskDatamodule (pl_datamodule.py:221-318)."""

    mask_mode = "channel"


class SyntheticTimeMaskDatamodule(SyntheticDatamodule):
    """Stand-in for HDF5TimeMaskDatamodule (pl_datamodule.py:320-421)."""

    mask_mode = "time"


def make_batch(system: str, batch: int, mask: str = "train", seed: int = 0, device="cpu") -> tuple:
    """One reference-shaped batch without a DataLoader (bench / smoke / tests)."""
    h, u = _FIELDS[system](batch, 128, first_seed=seed)
    h, u = torch.from_numpy(h), torch.from_numpy(u)
    res = h.shape[1]
    tg = torch.linspace(0, 1, res).reshape(1, res, 1, 1).expand(batch, res, res, 1).contiguous()
    xg = torch.linspace(0, 1, res).reshape(1, 1, res, 1).expand(batch, res, res, 1).contiguous()
    if mask == "train":
        m = torch.stack([sample_mask(h[i], u[i], True) for i in range(batch)])
    else:
        per = [sample_mask(h[i], u[i], False) for i in range(batch)]
        m = {k: torch.stack([p[k] for p in per]) for k in per[0]}
    mv = (lambda t: t.to(device)) if device != "cpu" else (lambda t: t)
    m = {k: mv(v) for k, v in m.items()} if isinstance(m, dict) else mv(m)
    return mv(h), mv(tg), mv(xg), mv(u), m


def field_stats(system: str, n: int = 64) -> Dict[str, torch.Tensor]:
    dm = SyntheticMaskDatamodule(system=system, n_train=n, n_test=1)
    return dm.get_norm_stats()
