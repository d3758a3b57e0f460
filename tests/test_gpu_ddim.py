"""GPU parity tests of BASELINE config 4: `PlDdim.sample_edm` (EDM sampler on the VP sigma grid with RePaint-style
conditioning, models/ddim.py:915-1051) on the kernels, against the oracle and the fixture of the unmodified reference
(tests/golden/ddim_path.pt).  Known-region kernels are bit-exact against the torch expressions; the per-evaluation
denoiser output is within the 16-bit bar (1e-2 relative) on the same input; observed entries are bit-identical."""
import copy

import pytest
import torch

from common import NoiseFeed, golden, hparams
from mcedm_b200 import data as D
from mcedm_b200.utils import randomize_zero_init, rel_l2
from oracle import edm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_repaint_kernels_bit_exact_against_torch(dev):
    from mcedm_b200 import _lib as L

    lib = L.lib()
    gen = torch.Generator().manual_seed(11)
    B, C, H, W = 2, 2, 128, 128
    hu, noise = torch.randn(B, C, H, W, generator=gen), torch.randn(B, C, H, W, generator=gen)
    mask = torch.ones(B, C, H, W)
    mask[:, 0, 0:, :] = 0.0
    mask[:, 1, 64:, :] = 0.0
    grid = O.VpGrid()
    t0 = grid.round_sigma(torch.tensor(80.0, dtype=torch.float64))
    aT = grid.compute_alpha(t0.long())
    known = hu * aT.sqrt() + noise * (1.0 - aT).sqrt()
    x0 = (known * mask + noise * (1.0 - mask)).to(torch.float64) * t0
    hu_d, noise_d, mask_d = [t.to(dev).contiguous() for t in (hu, noise, mask)]
    xg = torch.empty(B, C, H, W, device=dev, dtype=torch.float64)
    n, s = hu.numel(), L.stream_ptr()
    sa, s1 = float(aT.sqrt().reshape(())), float((1.0 - aT).sqrt().reshape(()))
    L.check(lib.mcedm_edm_vp_init(L.ptr(hu_d), L.ptr(noise_d), L.ptr(mask_d), sa, s1, float(t0), n, L.ptr(xg), s))
    assert torch.equal(xg.cpu(), x0)
    x_next = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64) * 3
    t_next = grid.round_sigma(torch.tensor(3.4, dtype=torch.float64))
    at = grid.compute_alpha(t_next.long())
    ref = (at.sqrt() * hu + (1 - at).sqrt() * noise) * mask + x_next * (1.0 - mask)
    xg = x_next.to(dev).contiguous()
    L.check(lib.mcedm_edm_repaint_blend(L.ptr(hu_d), L.ptr(noise_d), L.ptr(mask_d), float(at.sqrt().reshape(())),
                                        float((1 - at).sqrt().reshape(())), n, L.ptr(xg), s))
    assert torch.equal(xg.cpu(), ref)
    final = hu * mask + xg.cpu() * (1.0 - mask)
    L.check(lib.mcedm_edm_repaint_blend(L.ptr(hu_d), L.ptr(noise_d), L.ptr(mask_d), 1.0, 0.0, n, L.ptr(xg), s))
    assert torch.equal(xg.cpu(), final)
    assert torch.equal(xg.cpu()[mask == 1], hu.double()[mask == 1])


def _module(dev):
    from mcedm_b200.ddim import PlDdim

    cfg = hparams("config_adm_ddim_res32")
    torch.manual_seed(1)
    pl = PlDdim(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    sd = {k: v.detach().clone() for k, v in pl.model.state_dict().items()}
    return pl.to(dev).eval(), cfg, sd


def test_ddim_repaint_sampler_against_reference(dev):
    g = golden("ddim_path.pt")
    pl, cfg, sd = _module(dev)
    mcfg = dict(cfg.model.hparams.model)
    s = g["sample"]
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps, sp.n_time_h, sp.n_time_u, sp.n_repeat = s["steps"], s["n_time_h"], s["n_time_u"], s["n_repeat"]
    pl.set_test_sampler_params(sp)
    assert abs(pl.sigma_min - g["sigma_min"]) < 1e-12 and abs(pl.sigma_max - g["sigma_max"]) < 1e-9
    st = g["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    h, u = D._FIELDS["swe"](1, 128, first_seed=g["field_seed"])
    state = pl.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    feed = NoiseFeed(s["seed"])
    pl._noise_hook = feed.hook
    pl._trace = []
    xs = pl.sample_edm(state[..., :1], state[..., 1:2], sp, return_last=True, guide_dx=False)
    assert [tuple(c) for c in feed.calls] == [tuple(c) for c in s["calls"]]
    assert xs.shape == (1, 1, 128, 128, 2) and xs.dtype == torch.float64
    assert len(pl._trace) == len(s["denoised"])
    grid = O.VpGrid()
    for (i, k, which, sigma, d, xt), ref in zip(pl._trace, s["denoised"]):
        assert abs(sigma - ref["sigma"]) <= 1e-9 * max(1.0, ref["sigma"])          # identical sigma grid look-ups
        with torch.no_grad():
            d_or, _ = O.vp_denoise(sd, mcfg, grid, xt.cpu(), torch.tensor(sigma, dtype=torch.float64))
        assert rel_l2(d, d_or) < 1e-2
    # observed entries (u for t < 64) are the ground truth bit for bit; the generated part stays close over 10 evaluations
    assert torch.equal(xs[0, 0, :64, :, 1], state[0, :64, :, 1].double())
    assert torch.equal(xs[0, 0, :64, :, 1].cpu(), s["xs"][0, 0, :64, :, 1])
    assert rel_l2(xs, s["xs"]) < 5e-2
    pl._noise_hook, pl._trace = None, None
    # test_step: core metrics logged, known region error exactly zero (ddim.py:471-483 `test_u_known`)
    sp.n_samples = 2
    pl.set_test_sampler_params(sp)
    pl.set_pde_loss_function("swe", False)
    out = pl.test_step((torch.from_numpy(h).to(dev), None, None, torch.from_numpy(u).to(dev)), 0)
    assert float(pl.logged["test_u_known"]) == 0.0
    for k in ("test_mae_h", "test_mae_u", "test_mae_hu_un", "test_corr_h", "test_pde_loss", "test_pde_loss_gt"):
        assert torch.isfinite(torch.as_tensor(pl.logged[k])), k
    assert out["traj"].shape == (1, 1, 128, 128, 2, 2)


def test_vp_get_denoised_matches_oracle(dev):
    pl, cfg, sd = _module(dev)
    pl.set_test_sampler_params(cfg.diff_sampler)
    mcfg = dict(cfg.model.hparams.model)
    gen = torch.Generator().manual_seed(4)
    grid = O.VpGrid()
    for sig in (100.0, 2.0, 0.02):
        t = grid.round_sigma(torch.tensor(sig, dtype=torch.float64))
        xt = torch.randn(2, 2, 128, 128, generator=gen, dtype=torch.float64) * float(t)
        d, f = pl.get_denoised(pl.model, xt.to(dev), t)
        with torch.no_grad():
            d_or, f_or = O.vp_denoise(sd, mcfg, grid, xt, t)
        assert rel_l2(f, f_or) < 1e-2 and rel_l2(d, d_or) < 1e-2


# ------------------------------------------------------------------------------------------------ DDPM U-Net
def _ddpm_module(dev):
    """PlDdim as the reference ships config 4 (configs/config_ddim_res32.yaml: name 'ddim' -> ddim_blocks.Model) with the
    fixture's weights (tests/common.seeded_weights from the reference's state_dict shapes)."""
    from common import seeded_weights
    from mcedm_b200.ddim import PlDdim
    from mcedm_b200.ddpm_blocks import Model
    from mcedm_b200.utils import state_hash

    g = golden("ddpm_path.pt")
    cfg = hparams("config_ddim_res32")
    torch.manual_seed(1)
    pl = PlDdim(copy.deepcopy(cfg.model.hparams))
    assert isinstance(pl.model, Model)
    sd = seeded_weights(g["shapes"], seed=3)
    assert state_hash(sd) == g["weights_hash"]
    pl.model.load_state_dict(sd, strict=True)
    pl.ema_model.ma_model.load_state_dict(sd, strict=True)
    return pl.to(dev).eval(), cfg, sd, g


def test_ddpm_kernels_match_torch(dev):
    """csrc/ddpm.cu against the torch expressions of ddim_blocks.py: GroupNorm(32, eps 1e-6) of x + temb shift through
    per-channel statistics and folded coefficients (dense and padded-flat), decimation, timestep MLP."""
    import torch.nn.functional as F

    from mcedm_b200 import _lib as L

    lib = L.lib()
    st = L.stream_ptr()
    gen = torch.Generator().manual_seed(3)
    B, H, W = 3, 32, 32
    x = (torch.randn(B, H, W, 64, generator=gen) * 2 + 0.5).to(dev).half()
    gamma, beta = (1 + 0.1 * torch.randn(64, generator=gen)).to(dev), (0.1 * torch.randn(64, generator=gen)).to(dev)
    shift = (0.5 * torch.randn(B, 64, generator=gen)).to(dev)
    pitch, blk = C_geom(L, H, W)
    xf = torch.zeros(B, blk, 64, device=dev, dtype=torch.float16)
    xf[:, pitch:pitch + H * pitch].reshape(B, H, pitch, 64)[:, :, :W] = x
    for layout, src, npos in (("dense", x.contiguous(), H * W), ("flat", xf, blk)):
        for cpg, sh in ((2, shift), (4, None), (2, None)):
            part = torch.empty(B, 16, 64, 2, device=dev)
            coef = torch.empty(B, 128, device=dev)
            L.check(lib.mcedm_gn_stats16(L.ptr(src), npos, B, 1, 16, L.ptr(part), st))
            L.check(lib.mcedm_gn_coef_groups(L.ptr(part), 16, H * W, L.ptr(gamma), L.ptr(beta), cpg, 1e-6, L.ptr(sh),
                                             64 if sh is not None else 0, B, L.ptr(coef), st))
            xin = x.float().permute(0, 3, 1, 2) + (sh[:, :, None, None] if sh is not None else 0.0)
            ref = F.group_norm(xin, 64 // cpg, gamma, beta, eps=1e-6).permute(0, 2, 3, 1)
            got = coef[:, None, None, :64] * x.float() + coef[:, None, None, 64:]
            assert rel_l2(got, ref) < 2e-5, (layout, cpg, rel_l2(got, ref))
    # decimation: odd positions, dense -> flat and flat -> dense
    H2, W2 = 64, 64
    y = torch.randn(2, H2, W2, 64, generator=gen).to(dev).half().contiguous()
    p2, b2 = C_geom(L, H2 // 2, W2 // 2)
    out = torch.zeros(2, b2, 64, device=dev, dtype=torch.float16)
    L.check(lib.mcedm_decimate16(L.ptr(y), 0, 0, 2, H2, W2, L.ptr(out), p2, b2, st))
    got = out[:, p2:p2 + (H2 // 2) * p2].reshape(2, H2 // 2, p2, 64)[:, :, :W2 // 2]
    assert torch.equal(got, y[:, 1::2, 1::2])
    assert float(out.float().abs().sum()) == float(got.float().abs().sum())          # padding untouched
    pf, bf = C_geom(L, H2, W2)
    yf = torch.zeros(2, bf, 64, device=dev, dtype=torch.float16)
    yf[:, pf:pf + H2 * pf].reshape(2, H2, pf, 64)[:, :, :W2] = y
    out2 = torch.empty(2, H2 // 2, W2 // 2, 64, device=dev, dtype=torch.float16)
    L.check(lib.mcedm_decimate16(L.ptr(yf), pf, bf, 2, H2, W2, L.ptr(out2), 0, 0, st))
    assert torch.equal(out2, y[:, 1::2, 1::2])
    # timestep MLP + per-block projections
    from oracle.ddpm_oracle import timestep_embedding

    t = torch.tensor([999.0, 17.0, 0.0]).to(dev)
    w0, b0 = (torch.randn(256, 64, generator=gen) / 8).to(dev), (0.1 * torch.randn(256, generator=gen)).to(dev)
    w1, b1 = (torch.randn(256, 256, generator=gen) / 16).to(dev), (0.1 * torch.randn(256, generator=gen)).to(dev)
    wp, bp = (torch.randn(5, 64, 256, generator=gen) / 16).to(dev), (0.1 * torch.randn(5, 64, generator=gen)).to(dev)
    outt = torch.empty(5, 3, 64, device=dev)
    L.check(lib.mcedm_ddpm_temb(L.ptr(t), 3, L.ptr(w0), L.ptr(b0), L.ptr(w1), L.ptr(b1), L.ptr(wp), L.ptr(bp), 5, L.ptr(outt), st))
    sw = lambda v: v * torch.sigmoid(v)  # noqa: E731
    temb = F.linear(sw(F.linear(timestep_embedding(t.cpu(), 64).to(dev), w0, b0)), w1, b1)
    ref = torch.stack([F.linear(sw(temb), wp[i], bp[i]) for i in range(5)])
    assert rel_l2(outt, ref) < 1e-5, rel_l2(outt, ref)
    L.check_watchdog()


def C_geom(L, H, W):
    import ctypes as C

    p, b = C.c_int(0), C.c_int(0)
    L.check(L.lib().mcedm_flat_geometry(H, W, C.byref(p), C.byref(b)))
    return p.value, b.value


def test_ddpm_unet_forward_matches_reference(dev):
    """One evaluation of ddim_blocks.Model (per-sample timesteps, self-conditioning input zeros) on the launch plan against
    the output of the UNMODIFIED reference (tests/golden/make_golden_ddpm.py): 16-bit bar 1e-2."""
    from mcedm_b200 import _lib as L

    pl, cfg, sd, g = _ddpm_module(dev)
    gen = torch.Generator().manual_seed(g["forward"]["seed"])
    x = torch.randn(2, 2, 128, 128, generator=gen)
    with torch.no_grad():
        y = pl.model(x.to(dev), g["forward"]["t"].to(dev))
    L.check_watchdog()
    err = rel_l2(y, g["forward"]["y"])
    print("DDPM U-Net forward rel L2 vs reference:", err)
    assert err < 1e-2, err
    # rows are independent: a sub-batch alone is bit-identical (what makes row sharding exact for this network too)
    with torch.no_grad():
        y1 = pl.model(x[1:].to(dev).contiguous(), g["forward"]["t"][1:].to(dev))
    assert torch.equal(y1, y[1:])
    with pytest.raises(NotImplementedError):
        pl.model.train()(x.to(dev), g["forward"]["t"].to(dev))
    pl.model.eval()


def test_ddpm_config4_sampler_against_reference(dev):
    """BASELINE config 4 as shipped (configs/config_ddim_res32.yaml: PlDdim on the DDPM U-Net, edm_sampler with
    n_time_h=0, n_time_u=64, n_repeat=2): RNG call sequence, sigma look-ups, per-evaluation D_x within 1e-2 of the oracle
    on the same input, known region bit-identical, final sample close to the unmodified reference's."""
    from oracle import ddpm_oracle as DO

    pl, cfg, sd, g = _ddpm_module(dev)
    mcfg = g["model_cfg"]
    s = g["sample"]
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps, sp.n_time_h, sp.n_time_u, sp.n_repeat = s["steps"], s["n_time_h"], s["n_time_u"], s["n_repeat"]
    pl.set_test_sampler_params(sp)
    st = g["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    h, u = D._FIELDS["swe"](1, 128, first_seed=g["field_seed"])
    state = pl.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    feed = NoiseFeed(s["seed"])
    pl._noise_hook = feed.hook
    pl._trace = []
    xs = pl.sample_edm(state[..., :1], state[..., 1:2], sp, return_last=True, guide_dx=False)
    assert [tuple(c) for c in feed.calls] == [tuple(c) for c in s["calls"]]
    assert xs.shape == (1, 1, 128, 128, 2) and xs.dtype == torch.float64
    assert len(pl._trace) == len(s["denoised"])
    grid = O.VpGrid()
    worst = 0.0
    for (i, k, which, sigma, d, xt), ref in zip(pl._trace, s["denoised"]):
        assert abs(sigma - ref["sigma"]) <= 1e-9 * max(1.0, ref["sigma"])
        with torch.no_grad():
            d_or, _ = O.vp_denoise(sd, mcfg, grid, xt.cpu(), torch.tensor(sigma, dtype=torch.float64), net=DO.ddpm_net)
        worst = max(worst, rel_l2(d, d_or))
    print("DDPM config 4: worst per-evaluation D_x error", worst)
    assert worst < 1e-2, worst
    assert torch.equal(xs[0, 0, :64, :, 1], state[0, :64, :, 1].double())
    assert torch.equal(xs[0, 0, :64, :, 1].cpu(), s["xs"][0, 0, :64, :, 1])
    assert rel_l2(xs, s["xs"]) < 5e-2
    pl._noise_hook, pl._trace = None, None
    # graph replay path (use_cuda_graph) gives the same trajectory as eager launches
    feed2 = NoiseFeed(s["seed"])
    pl._noise_hook = feed2.hook
    pl.use_cuda_graph = False
    xs2 = pl.sample_edm(state[..., :1], state[..., 1:2], sp, return_last=True, guide_dx=False)
    assert torch.equal(xs2, xs)


def test_ddim_sample_with_repeat_against_reference(dev):
    """PlDdim.sample_with_repeat (models/ddim.py:808-913; the DDIM sampler with known-region replacement, repeats and
    self-conditioning) on the kernels with the DDPM U-Net: (1) the state-update kernels are bit-identical to the torch
    expressions on the traced inputs; (2) every network evaluation is within the 16-bit bar of the oracle on the SAME
    inputs (x_t, timestep vector, self-conditioning input); (3) the known region of every x0 prediction and of the final
    state is the data bit for bit; (4) RNG call sequence as the unmodified reference's."""
    from oracle import ddpm_oracle as DO

    pl, cfg, sd, gd = _ddpm_module(dev)
    g = golden("ddim_repeat.pt")
    mcfg = gd["model_cfg"]
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.type, sp.skip_type, sp.eta = "ddim", "uniform", 0.0
    sp.timesteps, sp.n_time_h, sp.n_time_u, sp.n_repeat = g["steps"], g["n_time_h"], g["n_time_u"], g["n_repeat"]
    st = g["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    pl.h_ch = pl.u_ch = 1
    h, u = D._FIELDS["swe"](1, 128, first_seed=g["field_seed"])
    state = pl.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    feed = NoiseFeed(g["seed"])
    pl._noise_hook = feed.hook
    pl._trace = []
    xs, x0 = pl.sample_with_repeat(state[..., :1], state[..., 1:2], sp, return_last=False)
    trace, pl._trace, pl._noise_hook = pl._trace, None, None
    assert [tuple(c) for c in feed.calls] == [tuple(c) for c in g["calls"]]
    assert xs.shape == g["xs"].shape and x0.shape == g["x0_preds"].shape and xs.dtype == torch.float32
    assert len(trace) == len(g["evals"])
    grid = O.VpGrid()
    hu = state.permute(0, 3, 1, 2).contiguous()
    mask = torch.ones_like(hu)
    mask[:, 0:1, g["n_time_h"]:, :] = 0.0
    mask[:, 1:2, g["n_time_u"]:, :] = 0.0
    worst = 0.0
    for rec, ref in zip(trace, g["evals"]):
        assert rec["t"] == ref["t"] and (rec["x_self_cond"] is not None) == ref["self_cond"]
        tv = torch.full((1,), rec["t"])
        with torch.no_grad():
            e_or = DO.ddpm_net(sd, mcfg, rec["xt"].cpu(), tv, None if rec["x_self_cond"] is None else rec["x_self_cond"].cpu())
        worst = max(worst, rel_l2(rec["et"], e_or))
    print("sample_with_repeat: worst per-evaluation e_t error", worst)
    assert worst < 1e-2, worst
    # state updates bit-identical to the torch expressions on the traced tensors (last timestep: t = 0, next = -1)
    rec = trace[-1]
    at, at_next = grid.compute_alpha(torch.tensor([0])).to(dev), grid.compute_alpha(torch.tensor([-1])).to(dev)
    x0_ref = (rec["xt"] - rec["et"] * (1 - at).sqrt()) / at.sqrt()
    x0_ref = hu * mask + x0_ref * (1.0 - mask)
    assert torch.equal(x0[0, -1].permute(2, 0, 1), x0_ref[0])
    c2 = (1 - at_next).sqrt()
    xn = at_next.sqrt() * x0_ref + c2 * rec["et"]
    hu_noise = NoiseFeed(g["seed"]).draw(hu.cpu()).to(dev)
    xn = (at_next.sqrt() * hu + c2 * hu_noise) * mask + xn * (1.0 - mask)
    assert torch.equal(xs[0, -1].permute(2, 0, 1), xn[0])
    for k in range(x0.shape[1]):
        assert torch.equal(x0[0, k, :64, :, 1], state[0, :64, :, 1])
    assert torch.equal(xs[0, -1, :64, :, 1], state[0, :64, :, 1])
    assert torch.isfinite(xs).all() and rel_l2(xs[:, 0], g["xs"][:, 0]) < 1e-6      # the initial state: same draws
