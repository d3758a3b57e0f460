"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container:  python tests/golden/make_golden.py
The fixtures pin the oracle (oracle/edm_oracle.py) and, through it, the CUDA path.  Everything that
can be regenerated deterministically (weights from seeds, synthetic fields, injected noise) is stored
only as a seed plus a sha256, so the fixtures stay small.

Conventions shared with the tests (tests/common.py):
  * weights: torch.manual_seed(1) -> reference constructor; zero-initialised tensors then replaced by
    mcedm_b200.utils.randomize_zero_init(seed=2) ("stress" weights; a fresh reference net outputs 0);
  * injected noise: every torch.randn_like call inside the reference is served from one CPU generator
    (seed given per fixture) as torch.randn(shape, dtype=like.dtype, generator=g).
"""
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
from mcedm_b200 import data as D  # noqa: E402
from mcedm_b200.utils import randomize_zero_init, state_hash  # noqa: E402


class NoiseFeed:
    """Serves randn_like draws from a seeded CPU generator; usable as a context manager that patches torch."""

    def __init__(self, seed):
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.calls = []

    def __call__(self, like, **kw):
        self.calls.append((tuple(like.shape), str(like.dtype)))
        return torch.randn(like.shape, dtype=like.dtype, generator=self.gen).to(like.device)

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = self
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


def normalized_state(system="swe_per", n=1, seed=0):
    """(h_unnorm, u_unnorm, stats) as [n,128,128,1] tensors plus the datamodule statistics."""
    h, u = D._FIELDS[system](n, 128, first_seed=seed)
    st = D.field_stats(system, 16)
    return torch.from_numpy(h), torch.from_numpy(u), st


def build_reference_module(ref, config_name):
    cfg = R.reference_hparams(config_name)
    hp = cfg.model.hparams
    torch.manual_seed(1)
    cls = ref.mcedm.PlMcedm if config_name == "config_adm_edm_mcedm_res32" else ref.ddim.PlCondEdm
    pl = cls(copy.deepcopy(hp))
    init_hash = state_hash(pl.model.state_dict())
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    return pl, hp, cfg, init_hash


def main():
    torch.set_num_threads(8)
    ref = R.import_reference()
    out = {}

    # ------------------------------------------------------------------ G1: U-Net forward
    pl, hp, cfg, init_hash = build_reference_module(ref, "config_adm_edm_mcedm_res32")
    net = pl.model
    g = torch.Generator().manual_seed(10)
    cases = []
    for B, labels in [(1, [0.3]), (2, [-1.0, 0.8])]:
        x = torch.randn(B, 2, 128, 128, generator=g)
        c = torch.randn(B, 2, 128, 128, generator=g)
        nl = torch.tensor(labels)
        with torch.no_grad():
            y = net(x, nl, c)
        cases.append(dict(x=x, cond=c, noise_labels=nl, out=y))
    torch.save(dict(init_hash=init_hash, stress_hash=state_hash(net.state_dict()), cases=cases,
                    n_params=sum(p.numel() for p in net.parameters())), os.path.join(HERE, "unet_forward.pt"))
    print("G1 done", init_hash[:12], cases[0]["out"].abs().max().item())

    # ------------------------------------------------------------------ G2: get_denoised at several sigmas
    h, u, stats = normalized_state("swe_per", 1, seed=7)
    pl.normalizer_input.set_stats(stats["input_mean"], stats["input_std"])
    pl.normalizer_target.set_stats(stats["target_mean"], stats["target_std"])
    pl.h_ch = pl.u_ch = 1
    state = pl.data_transform(h, u)                                   # b h w c
    state_c = state.permute(0, 3, 1, 2).contiguous()
    masks = ref.h5_dataset.HDF5MaskDataset.sample_mask(type("S", (), {"is_train": False})(), h[0], u[0])
    den = []
    gg = torch.Generator().manual_seed(11)
    for sigma in (80.0, 1.5, 0.05):
        m = masks["u"].unsqueeze(0).permute(0, 3, 1, 2)
        cond = state_c * (1 - m) + torch.randn(state_c.shape, generator=gg) * m
        xt = (state_c + torch.randn(state_c.shape, generator=gg) * sigma * m).double()
        with torch.no_grad():
            d, f = pl.get_denoised(pl.ema_model, xt, torch.tensor(sigma, dtype=torch.float64), cond=cond, w=0.0)
        den.append(dict(sigma=sigma, xt=xt, cond=cond, D=d, F=f))
    torch.save(dict(cases=den), os.path.join(HERE, "denoise.pt"))
    print("G2 done")

    # ------------------------------------------------------------------ G3: short trajectories with injected noise
    trajs = []
    time_masks = ref.h5_dataset.HDF5TimeMaskDataset.sample_mask(
        type("S", (), {"is_train": False, "add_time_masks": True})(), h[0], u[0])
    for name, mask_hwc, steps, seed in [("u", masks["u"], 4, 21), ("h_time", time_masks["h"], 3, 22)]:
        sp = copy.deepcopy(cfg.diff_sampler)
        sp.timesteps = steps
        mask = mask_hwc.unsqueeze(0)
        rec = []
        orig = pl.get_denoised

        def traced(model, xt, t, **kw):
            d, f = orig(model, xt, t, **kw)
            rec.append(dict(sigma=float(t), D=d.clone()))
            return d, f

        pl.get_denoised = traced
        with NoiseFeed(seed) as feed, torch.no_grad():
            cond_in = pl.get_cond_in(state, mask, None, None).permute(0, 3, 1, 2).contiguous()
            mask_c = mask.permute(0, 3, 1, 2).contiguous()
            noise = torch.randn_like(state_c)
            xs = pl.sample_edm(noise, cond_in, mask_c, sp, return_last=True, guide_dx=False)
        pl.get_denoised = orig
        trajs.append(dict(name=name, steps=steps, seed=seed, mask=mask, state=state, denoised=rec, xs=xs,
                          calls=feed.calls))
        print("G3", name, len(rec), xs.shape, xs.dtype)
    torch.save(dict(trajs=trajs, stats={k: v for k, v in stats.items() if torch.is_tensor(v)}),
               os.path.join(HERE, "trajectory.pt"))

    # ------------------------------------------------------------------ G4: training step (loss + a few gradients)
    B = 2
    hb, ub, _ = normalized_state("swe_per", B, seed=30)
    torch.manual_seed(123)
    mb = torch.stack([ref.h5_dataset.HDF5MaskDataset.sample_mask(type("S", (), {"is_train": True})(), hb[i], ub[i])
                      for i in range(B)])
    grid = torch.zeros(B, 128, 128, 1)
    with NoiseFeed(31):
        torch.manual_seed(5)     # CPU RNG for sigma (:269) and the cond-drop draw (:231)
        pl.zero_grad()
        loss = pl.training_step((hb, grid, grid, ub, mb), 0)
        loss.backward()
    grads = {k: p.grad.clone() for k, p in pl.model.named_parameters()
             if k in ("enc.128x128_conv.weight", "dec.32x32_block0.qkv.weight", "out_conv.weight",
                      "enc.64x64_down.affine.weight", "dec.128x128_block1.skip.weight", "map_layer0.weight",
                      "dec.64x64_up.norm1.weight", "enc.32x32_block0.proj.bias")}
    gnorm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in pl.model.parameters() if p.grad is not None))
    torch.save(dict(seed_fields=30, mask=mb, loss=loss.detach(), grads=grads, grad_norm=gnorm, noise_seed=31,
                    cpu_seed=5), os.path.join(HERE, "train_step.pt"))
    print("G4 done", float(loss), float(gnorm))

    # ------------------------------------------------------------------ G5: masks
    torch.manual_seed(77)
    S = type("S", (), {"is_train": True})()
    train_masks = torch.stack([ref.h5_dataset.HDF5MaskDataset.sample_mask(S, h[0], u[0]) for _ in range(8)])
    torch.manual_seed(78)
    ST = type("S", (), {"is_train": True, "add_time_masks": False})()
    ST.get_train_mask = lambda a, b: ref.h5_dataset.HDF5TimeMaskDataset.get_train_mask(ST, a, b)
    time_train = torch.stack([ref.h5_dataset.HDF5TimeMaskDataset.sample_mask(ST, h[0], u[0]) for _ in range(8)])
    torch.save(dict(train_seed=77, train_masks=train_masks.to(torch.uint8), time_seed=78,
                    time_train=time_train.to(torch.uint8),
                    eval_u=masks["u"].to(torch.uint8), eval_h=masks["h"].to(torch.uint8),
                    time_eval={k: v.to(torch.uint8) for k, v in time_masks.items()}),
               os.path.join(HERE, "masks.pt"))
    print("G5 done")

    # ------------------------------------------------------------------ G6: config 5 (PlCondEdm) network
    pl5, hp5, cfg5, init5 = build_reference_module(ref, "config_adm_edm_res32_cond_h")
    g = torch.Generator().manual_seed(40)
    x = torch.randn(2, 1, 128, 128, generator=g)
    c = torch.randn(2, 1, 128, 128, generator=g)
    nl = torch.tensor([0.2, -0.9])
    with torch.no_grad():
        y = pl5.model(x, nl, c)
    torch.save(dict(init_hash=init5, stress_hash=state_hash(pl5.model.state_dict()), x=x, cond=c, noise_labels=nl,
                    out=y), os.path.join(HERE, "cond_edm_forward.pt"))
    print("G6 done", y.shape)

    # ------------------------------------------------------------------ G7: full 50-step trajectory (final state only)
    sp = copy.deepcopy(cfg.diff_sampler)
    mask = masks["h"].unsqueeze(0)
    with NoiseFeed(50) as feed, torch.no_grad():
        cond_in = pl.get_cond_in(state, mask, None, None).permute(0, 3, 1, 2).contiguous()
        mask_c = mask.permute(0, 3, 1, 2).contiguous()
        noise = torch.randn_like(state_c)
        xs = pl.sample_edm(noise, cond_in, mask_c, sp, return_last=True, guide_dx=False)
    rmse = torch.sqrt((((xs[:, -1] - state) * mask) ** 2).sum() / mask.sum())
    torch.save(dict(seed=50, mask_name="h", xs=xs, rmse=rmse, n_calls=len(feed.calls)),
               os.path.join(HERE, "trajectory_full.pt"))
    print("G7 done rmse", float(rmse))


if __name__ == "__main__":
    main()
