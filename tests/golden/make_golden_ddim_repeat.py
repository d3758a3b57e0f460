"""Generates tests/golden/ddim_repeat.pt by running the UNMODIFIED reference `PlDdim.sample_with_repeat`
(models/ddim.py:808-913: DDIM sampler on the VP schedule with RePaint-style known-region replacement, n_repeat
re-evaluations per timestep, self-conditioning on the previous x0 prediction) on CPU with the DDPM U-Net the reference
builds for `name: ddim`:   python tests/golden/make_golden_ddim_repeat.py

Same weights as ddpm_path.pt (tests/common.seeded_weights from the state_dict shapes), dam-break-shaped fields, injected
noise.  Stored: the final xs / x0_preds, and for every network evaluation its inputs' hashes are NOT stored — the tests
recompute them; only (t, the network output e_t) per evaluation so the oracle can be pinned step by step."""
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

sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import seeded_weights  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref = R.import_reference()
    cfg = R.reference_hparams("config_ddim_res32")
    hp = cfg.model.hparams
    torch.manual_seed(1)
    pl = ref.ddim.PlDdim(copy.deepcopy(hp))
    shapes = {k: tuple(v.shape) for k, v in pl.model.state_dict().items()}
    pl.model.load_state_dict(seeded_weights(shapes, seed=3))
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.type, sp.skip_type, sp.eta = "ddim", "uniform", 0.0
    sp.timesteps, sp.n_time_h, sp.n_time_u, sp.n_repeat = 3, 0, 64, 2
    h, u, stats = normalized_state("swe", 1, seed=93)
    pl.normalizer_input.set_stats(stats["input_mean"], stats["input_std"])
    pl.normalizer_target.set_stats(stats["target_mean"], stats["target_std"])
    pl.h_ch = pl.u_ch = 1
    state = pl.data_transform(h, u)
    rec = []
    net = pl.ema_model.ma_model
    orig = net.forward

    def traced(x, t, *a, **kw):
        y = orig(x, t, *a, **kw)
        rec.append(dict(t=float(t.reshape(-1)[0]), et=y.detach().clone(),
                        self_cond=kw.get("x_self_cond") is not None))
        return y

    net.forward = traced
    with NoiseFeed(94) as feed, torch.no_grad():
        xs, x0 = pl.sample_with_repeat(state[..., :1], state[..., 1:2], sp, return_last=False, guide_dx=False)
    print("sample_with_repeat", xs.shape, x0.shape, len(rec), [r["t"] for r in rec], feed.calls)
    torch.save(dict(field_seed=93, seed=94, steps=3, n_time_h=0, n_time_u=64, n_repeat=2, evals=rec, xs=xs, x0_preds=x0,
                    calls=feed.calls, stats={k: v for k, v in stats.items() if torch.is_tensor(v)}),
               os.path.join(HERE, "ddim_repeat.pt"))


if __name__ == "__main__":
    main()
