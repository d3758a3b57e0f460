"""Helpers shared by the CPU and GPU tests."""
import copy
import os

import torch

from mcedm_b200 import data as D
from mcedm_b200.config import compose
from mcedm_b200.utils import randomize_zero_init, state_hash

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def hparams(config_name="config_adm_edm_mcedm_res32"):
    return compose(config_name)


def stress_unet(config_name="config_adm_edm_mcedm_res32"):
    """mcedm_b200.DhariwalUNet with the fixture weights: seed 1 init, zero-init tensors randomised (seed 2)."""
    from mcedm_b200.adm_blocks import DhariwalUNet

    cfg = hparams(config_name)
    torch.manual_seed(1)
    net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
    init_hash = state_hash(net.state_dict())
    randomize_zero_init(net, 2)
    return net, cfg, init_hash


def stress_module(config_name="config_adm_edm_mcedm_res32"):
    """mcedm_b200.PlMcedm with the fixture weights in both the live and the EMA network."""
    from mcedm_b200.mcedm import PlMcedm

    cfg = hparams(config_name)
    torch.manual_seed(1)
    pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    return pl, cfg


class NoiseFeed:
    """Serves the reference's randn_like draws from a seeded CPU generator (same protocol as
    tests/golden/make_golden.py): torch.randn(shape, dtype=like.dtype, generator=g)."""

    def __init__(self, seed):
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.calls = []

    def draw(self, like):
        self.calls.append((tuple(like.shape), str(like.dtype)))
        return torch.randn(like.shape, dtype=like.dtype, generator=self.gen).to(like.device)

    def hook(self, kind, like):
        return self.draw(like)


def fixture_state(system="swe_per", n=1, seed=7):
    """Normalised (state b h w c, h_unnorm, u_unnorm, stats) exactly as make_golden.normalized_state + data_transform."""
    h, u = D._FIELDS[system](n, 128, first_seed=seed)
    h, u = torch.from_numpy(h), torch.from_numpy(u)
    st = D.field_stats(system, 16)
    state = torch.cat([(h - st["input_mean"]) / st["input_std"], (u - st["target_mean"]) / st["target_std"]], dim=-1)
    return state, h, u, st


def pde_fields(system, n, seed, amp=0.05):
    """(pred, gt, stats) of tests/golden/make_golden_pde.perturbed_fields: clean synthetic field + seeded perturbation."""
    h, u = D._FIELDS[system](n, 128, first_seed=seed)
    gt = torch.cat([torch.from_numpy(h), torch.from_numpy(u)], dim=-1)
    g = torch.Generator().manual_seed(1000 + seed)
    pred = gt + amp * torch.randn(gt.shape, generator=g)
    return pred, gt, D.field_stats(system, 16)


def seeded_weights(shapes, seed=3):
    """Deterministic fp32 weights for a network given only its state_dict shapes (used for the DDPM U-Net fixture, whose
    module has no mirror in this repo yet): N(0, 1/fan_in) for matrices / filters, 1 + 0.1 N for norm scales
    (`norm*.weight`), 0.1 N for every other vector.  One CPU generator, tensors drawn in the given order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            out[name] = torch.randn(shape, generator=g) / fan_in ** 0.5
        elif "norm" in name and name.endswith("weight"):
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            out[name] = 0.1 * torch.randn(shape, generator=g)
    return out
