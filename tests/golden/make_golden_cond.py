"""Generates tests/golden/cond_edm_path.pt by running the UNMODIFIED reference PlCondEdm (config 5,
config_adm_edm_res32_cond_h: u denoised, h as condition, no mask) on CPU:  python tests/golden/make_golden_cond.py

  * a 3-step PlCondDdim.sample_edm trajectory (models/ddim.py:1532-1601) on Darcy-shaped fields with injected noise:
    per-evaluation D_x, final xs, the RNG call sequence;
  * one PlCondEdm.training_step (models/ddim.py:1700-1737): loss and gradient norm.
Same conventions as make_golden.py (stress weights, NoiseFeed)."""
from __future__ import annotations

import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as R  # noqa: E402
from make_golden import NoiseFeed, build_reference_module, normalized_state  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref = R.import_reference()
    pl, hp, cfg, init_hash = build_reference_module(ref, "config_adm_edm_res32_cond_h")
    a, u, stats = normalized_state("darcy", 2, seed=60)
    pl.normalizer_input.set_stats(stats["input_mean"], stats["input_std"])
    pl.normalizer_target.set_stats(stats["target_mean"], stats["target_std"])
    pl.h_ch = pl.u_ch = 1
    pl.cond_p = 1.0                         # keep the conditioning on (PlCondDdim default 0.8 drops it at random)

    # ---- sampling
    state = pl.data_transform(a[:1], u[:1])
    h_n, u_n = state[..., :1], state[..., 1:2]
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps = 3
    rec = []
    orig = pl.get_denoised

    def traced(model, xt, t, **kw):
        d, f = orig(model, xt, t, **kw)
        rec.append(dict(sigma=float(t), D=d.clone()))
        return d, f

    pl.get_denoised = traced
    with NoiseFeed(61) as feed, torch.no_grad():
        cond_in = pl.get_cond_in(h_n, u_n, None, None)
        u_noise = torch.randn_like(u_n)
        xs = pl.sample_edm(cond_in, u_noise, sp, return_last=True, guide_dx=False)
    pl.get_denoised = orig
    print("sample", len(rec), xs.shape, xs.dtype, feed.calls)

    # ---- training step
    grid = torch.zeros(2, 128, 128, 1)
    with NoiseFeed(62):
        torch.manual_seed(6)
        pl.zero_grad()
        loss = pl.training_step((a, grid, grid, u), 0)
        loss.backward()
    gnorm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in pl.model.parameters() if p.grad is not None))
    print("train", float(loss), float(gnorm))
    torch.save(dict(field_seed=60, sample=dict(steps=3, seed=61, denoised=rec, xs=xs, calls=feed.calls),
                    train=dict(noise_seed=62, cpu_seed=6, loss=loss.detach(), grad_norm=gnorm),
                    stats={k: v for k, v in stats.items() if torch.is_tensor(v)}), os.path.join(HERE, "cond_edm_path.pt"))


if __name__ == "__main__":
    main()
