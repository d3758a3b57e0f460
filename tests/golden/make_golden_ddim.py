"""Generates tests/golden/ddim_path.pt by running the UNMODIFIED reference `PlDdim` (models/ddim.py) on CPU:
python tests/golden/make_golden_ddim.py

BASELINE config 4: RePaint-conditioned EDM sampling (`diff_sampler=edm_sampler`, n_time_h=0, n_time_u=64, n_repeat=2)
on dam-break-shaped SWE fields.  The module is built with `hparams.name = "adm_ddim"`, for which the reference
constructs `DhariwalUNet` (models/ddim.py:40-41); hparams = mcedm_b200/configs/model/adm_ddim_res32.yaml (the
reference's ddim_res32.yaml values plus the three keys DhariwalUNet reads).  A 3-step trajectory with injected noise:
per-evaluation D_x, final xs, the RNG call sequence, the VP grid scalars.  Same conventions as make_golden.py."""
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
from make_golden import NoiseFeed, normalized_state  # noqa: E402
from mcedm_b200.config import compose  # noqa: E402
from mcedm_b200.utils import randomize_zero_init, state_hash  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref = R.import_reference()
    cfg = compose("config_adm_ddim_res32")
    torch.manual_seed(1)
    pl = ref.ddim.PlDdim(copy.deepcopy(cfg.model.hparams))
    assert type(pl.model).__name__ == "DhariwalUNet"
    init_hash = state_hash(pl.model.state_dict())
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps, sp.n_time_h, sp.n_time_u, sp.n_repeat = 3, 0, 64, 2
    pl.set_test_sampler_params(sp)
    h, u, stats = normalized_state("swe", 1, seed=80)
    pl.normalizer_input.set_stats(stats["input_mean"], stats["input_std"])
    pl.normalizer_target.set_stats(stats["target_mean"], stats["target_std"])
    pl.h_ch = pl.u_ch = 1
    state = pl.data_transform(h, u)
    rec = []
    orig = pl.get_denoised

    def traced(model, xt, t, **kw):
        d, f = orig(model, xt, t, **kw)
        rec.append(dict(sigma=float(t), D=d.clone()))
        return d, f

    pl.get_denoised = traced
    with NoiseFeed(81) as feed, torch.no_grad():
        xs = pl.sample_edm(state[..., :1], state[..., 1:2], sp, return_last=True, guide_dx=False)
    pl.get_denoised = orig
    known = (xs[0, 0, :64, :, 1] - state[0, :64, :, 1].double()).abs().max()
    print("ddim sample", len(rec), xs.shape, xs.dtype, feed.calls, "known-region error", float(known),
          [r["sigma"] for r in rec])
    torch.save(dict(field_seed=80, init_hash=init_hash, sigma_min=pl.sigma_min, sigma_max=pl.sigma_max,
                    sample=dict(steps=3, n_time_h=0, n_time_u=64, n_repeat=2, seed=81, denoised=rec, xs=xs,
                                calls=feed.calls),
                    stats={k: v for k, v in stats.items() if torch.is_tensor(v)}), os.path.join(HERE, "ddim_path.pt"))


if __name__ == "__main__":
    main()
