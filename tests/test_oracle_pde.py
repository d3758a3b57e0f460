"""CPU: the PDE-residual oracle (oracle/pde_oracle.py, numpy float32) against tests/golden/pde.pt, the outputs of the
unmodified reference (tests/golden/make_golden_pde.py).  Residual matrices are bit-exact; gradients (autograd in the
reference, analytic adjoint here) within 1e-5 of their maximum."""
import copy

import numpy as np
import pytest
import torch

from common import NoiseFeed, golden, pde_fields, stress_unet
from mcedm_b200 import data as D
from mcedm_b200.utils import rel_l2
from oracle import edm_oracle as O
from oracle import pde_oracle as P


def _fl(st):
    return {k: float(v) for k, v in st.items() if torch.is_tensor(v)}


@pytest.mark.parametrize("system", ["swe_per", "swe"])
def test_swe_residual_matrix_bit_exact_and_gradient(system):
    g = golden("pde.pt")[system]
    pred, gt, st = pde_fields(system, 2, g["field_seed"])
    s = _fl(g["stats"])
    assert s == _fl(st)
    _, _, _, dt, dx = P.pde_params(system, 128, 128)
    for name, target in (("self", pred), ("gt", gt)):
        m = P.swe_fv_loss_matrix(pred.numpy(), target.numpy(), s["input_std"], s["target_std"], dt, dx)
        assert np.array_equal(m, g[f"loss_{name}"].numpy())
        gr = P.swe_fv_grad(pred.numpy(), target.numpy(), s["input_std"], s["target_std"], dt, dx)
        ref = g[f"grad_{name}"].numpy()
        assert np.abs(gr - ref).max() < 1e-5 * np.abs(ref).max()
    m = P.swe_fv_loss_matrix(pred.numpy(), gt.numpy(), s["input_std"], s["target_std"], dt, dx)
    assert np.array_equal(np.minimum(m, np.float32(1.0)), g["loss_gt_clamped"].numpy())


def test_darcy_residual_matrix_bit_exact():
    g = golden("pde.pt")["darcy"]
    a, u = D._FIELDS["darcy"](2, 128, first_seed=g["field_seed"])
    m = P.darcy_loss_matrix(np.concatenate([a, u], axis=-1))
    assert m.shape == (2, 124, 124)
    assert np.array_equal(m, g["loss"].numpy())


def test_module_level_pde_loss_and_guidance_gradient():
    g = golden("pde.pt")
    # PlMcedm.get_pde_loss: float64 normalised sample b h w c
    gm = g["mcedm"]
    s = _fl(gm["stats"])
    h, u = D._FIELDS["swe_per"](2, 128, first_seed=gm["field_seed"])
    state = torch.cat([(torch.from_numpy(h) - s["input_mean"]) / s["input_std"],
                       (torch.from_numpy(u) - s["target_mean"]) / s["target_std"]], dim=-1)
    gen = torch.Generator().manual_seed(gm["noise_seed"])
    sample = state.double() + 0.1 * torch.randn(state.shape, generator=gen, dtype=torch.float64)
    _, tot = P.get_pde_loss(sample[..., 0].numpy(), sample[..., 1].numpy(), s, "swe_per")
    assert abs(tot - float(gm["pde_sample"])) < 1e-5 * float(gm["pde_sample"])
    _, tot = P.get_pde_loss(state[..., 0].numpy(), state[..., 1].numpy(), s, "swe_per")
    assert abs(tot - float(gm["pde_gt"])) < 1e-5 * float(gm["pde_gt"])
    # PlCondEdm.get_pde_loss / get_dx_pde
    gc = g["cond"]
    s = _fl(gc["stats"])
    h, u = D._FIELDS["swe_per"](1, 128, first_seed=gc["field_seed"])
    h_n = (torch.from_numpy(h) - s["input_mean"]) / s["input_std"]
    u_n = (torch.from_numpy(u) - s["target_mean"]) / s["target_std"]
    gen = torch.Generator().manual_seed(gc["noise_seed"])
    u_s = u_n.double() + 0.1 * torch.randn(u_n.shape, generator=gen, dtype=torch.float64)
    _, tot = P.get_pde_loss(h_n[..., 0].numpy(), u_s[..., 0].numpy(), s, "swe_per")
    assert abs(tot - float(gc["pde"])) < 1e-5 * float(gc["pde"])
    for calc_prob, key in ((True, "dx_mean"), (False, "dx_sum")):
        d = P.get_dx_pde_cond(h_n[..., 0].numpy(), u_s[..., 0].numpy(), s, "swe_per", calc_prob=calc_prob)
        ref = gc[key].numpy()
        assert d.shape == ref.shape
        assert np.abs(d - ref).max() < 1e-5 * np.abs(ref).max()


def test_guided_cond_sampler_matches_reference():
    """3-step PDE-guided PlCondDdim.sample_edm (guide_dx=True): oracle sampler + oracle gradient vs the reference."""
    gc = golden("pde.pt")["cond"]
    s = _fl(gc["stats"])
    net, cfg, _ = stress_unet("config_adm_edm_res32_cond_h")
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    mcfg = dict(cfg.model.hparams.model)
    h, u = D._FIELDS["swe_per"](1, 128, first_seed=gc["field_seed"])
    h_n = ((torch.from_numpy(h) - s["input_mean"]) / s["input_std"]).permute(0, 3, 1, 2).contiguous()
    sp = dict(copy.deepcopy(cfg.diff_sampler))
    sp["timesteps"] = gc["sample"]["steps"]

    def guide(hc, den):
        return torch.from_numpy(P.get_dx_pde_cond(hc[:, 0].numpy(), den[:, 0].numpy(), s, "swe_per", calc_prob=True))

    for key, gfn in (("xs", guide), ("xs_plain", None)):
        feed = NoiseFeed(gc["sample"]["seed"])
        u_noise = feed.draw(torch.empty(1, 128, 128, 1)).permute(0, 3, 1, 2).contiguous()
        rec = []
        with torch.no_grad():
            xs = O.cond_sample_edm(sd, mcfg, u_noise, h_n, sp, lambda i, x: feed.draw(x), record=rec, guide=gfn)
        if gfn is not None:
            assert [tuple(c) for c in feed.calls] == [tuple(c) for c in gc["sample"]["calls"]]
            for (i, which, sigma, d), ref in zip(rec, gc["sample"]["denoised"]):
                assert rel_l2(d, ref["D"]) < 1e-5
        assert rel_l2(xs, gc["sample"][key]) < 1e-5
    # the guidance term is not a no-op on this fixture
    assert float((gc["sample"]["xs"] - gc["sample"]["xs_plain"]).abs().max()) > 1e-2
