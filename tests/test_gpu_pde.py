"""GPU parity tests of the PDE-residual path (K6, csrc/pde.cu) and the PDE-guided sampler, through the C ABI and the
reference-shaped Python surface, against the numpy oracle and the fixture of the unmodified reference
(tests/golden/pde.pt).

Tolerances: residual matrices bit-exact (float32, every operation in torch's order); sums 1e-6 relative (float64 fixed
order here, float32 torch.sum in the reference); gradients 2e-5 of their maximum (analytic float32 adjoint here,
float32 autograd in the reference); guided fp64 updates bit-exact given the same D and gradient."""
import copy

import numpy as np
import pytest
import torch

from common import NoiseFeed, golden, hparams, pde_fields
from mcedm_b200 import data as D
from mcedm_b200.utils import rel_l2
from oracle import edm_oracle as O
from oracle import pde_oracle as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _normalizers(st, dev):
    from mcedm_b200.nn_misc import Normalizer

    return (Normalizer(st["input_mean"].to(dev), st["input_std"].to(dev)),
            Normalizer(st["target_mean"].to(dev), st["target_std"].to(dev)))


@pytest.mark.parametrize("system", ["swe_per", "swe"])
def test_swe_residual_and_gradient_kernels(dev, system):
    from mcedm_b200.pde_loss import get_pde_loss_function

    g = golden("pde.pt")[system]
    pred, gt, st = pde_fields(system, 2, g["field_seed"])
    nh, nu = _normalizers(st, dev)
    f, _ = get_pde_loss_function(system, False)
    pd, gd = pred.to(dev), gt.to(dev)
    # module contract: forward(pred, gt, ...) on un-normalised [B,T,X,2]
    assert torch.equal(f(pd, pd, nh, nu).cpu(), g["loss_self"])
    assert torch.equal(f(pd, gd, nh, nu).cpu(), g["loss_gt"])
    assert torch.equal(f(pd, gd, nh, nu, clamp_loss=True).cpu(), g["loss_gt_clamped"])
    for name, target in (("self", pd), ("gt", gd)):
        gr = f(pd, target, nh, nu, return_d=True).cpu()
        ref = g[f"grad_{name}"]
        assert gr.shape == ref.shape
        assert float((gr - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    # fused entry: normalised float64 planes with NCHW strides, inverse normalisation inside the kernel
    s = {k: float(v) for k, v in st.items() if torch.is_tensor(v)}
    state = torch.cat([(pred[..., :1] - s["input_mean"]) / s["input_std"],
                       (pred[..., 1:] - s["target_mean"]) / s["target_std"]], dim=-1).double()
    x = state.permute(0, 3, 1, 2).contiguous().to(dev)                # b c h w float64
    m, tot = f.residual(x[:, 0], x[:, 1], nh, nu, want_matrix=True)
    m_or, tot_or = P.get_pde_loss(state[..., 0].numpy(), state[..., 1].numpy(), s, system)
    assert np.array_equal(m.cpu().numpy(), m_or)
    assert abs(float(tot) - tot_or) <= 1e-9 * tot_or
    assert abs(float(tot) - float(m.double().sum())) <= 1e-12 * tot_or
    # channel-mean / channel-sum output modes of the gradient
    g0 = f.gradient(x[:, 0], x[:, 1], nh, nu, mode=0)
    assert torch.equal(f.gradient(x[:, 0], x[:, 1], nh, nu, mode=1), (g0[..., 0] + g0[..., 1]) / 2)
    assert torch.equal(f.gradient(x[:, 0], x[:, 1], nh, nu, mode=2), g0[..., 0] + g0[..., 1])


def test_swe_residual_edge_cases(dev):
    """NaN handling (pde_loss.py:211, :241), dry cells (h = 0), non-square and odd sizes, against the oracle."""
    from mcedm_b200.nn_misc import Normalizer
    from mcedm_b200.pde_loss import SweFvLoss

    nh, nu = Normalizer(torch.tensor(0.0), torch.tensor(1.5)).to(dev), Normalizer(torch.tensor(0.0), torch.tensor(0.7)).to(dev)
    gen = torch.Generator().manual_seed(3)
    for B, T, X in ((1, 2, 2), (3, 7, 33), (2, 5, 1020), (1, 128, 128)):
        f = SweFvLoss(Tn=0.2 * T / X, x_min=-0.5, x_max=0.5)          # dt = 0.2 dx: a stable step at every size
        h = 1.0 + torch.rand(B, T, X, generator=gen)
        u = torch.randn(B, T, X, generator=gen) * 0.3
        if X > 8:
            h[0, 1, 3] = float("nan")
            h[0, 0, 5] = 0.0
            u[0, 0, 5] = 0.0
        pred = torch.stack([h, u], dim=-1)
        step = (f.x_max - f.x_min) / X
        xg = f.gen_x(X)
        dt, dx = f.Tn / T, float(xg[1] - xg[0])
        m_or = P.swe_fv_loss_matrix(pred.numpy(), pred.numpy(), 1.5, 0.7, dt, dx)
        m = f(pred.to(dev), pred.to(dev), nh, nu).cpu().numpy()
        assert np.array_equal(m, m_or, equal_nan=True), (B, T, X)
        if X > 8:                                                     # the dry cell makes the adjoint ill-conditioned
            pred[0, 0, 5, 0] = 1.0                                    # (1/(h+1e-8)^2 terms): gradient without it
        g_or = P.swe_fv_grad(pred.numpy(), pred.numpy(), 1.5, 0.7, dt, dx)
        gk = f(pred.to(dev), pred.to(dev), nh, nu, return_d=True).cpu().numpy()
        assert not np.isnan(gk).any()
        assert np.abs(gk - g_or).max() <= 2e-5 * max(np.abs(g_or).max(), 1e-30), (B, T, X)


def test_darcy_residual_kernel(dev):
    from mcedm_b200.pde_loss import get_pde_loss_function

    g = golden("pde.pt")["darcy"]
    a, u = D._FIELDS["darcy"](2, 128, first_seed=g["field_seed"])
    st = D.field_stats("darcy", 16)
    nh, nu = _normalizers(st, dev)
    f, _ = get_pde_loss_function("darcy", False)
    x = torch.cat([torch.from_numpy(a), torch.from_numpy(u)], dim=-1).to(dev)
    m = f(x, x, nh, nu)
    assert torch.equal(m.cpu(), g["loss"])
    _, tot = f.residual(x[..., 0], x[..., 1], nh, nu, apply_norm=False)
    assert abs(float(tot) - float(g["loss"].double().sum())) <= 1e-9 * float(tot)
    with pytest.raises(NotImplementedError):
        f(x, x, nh, nu, return_d=True)


def test_guided_update_kernels_bit_exact_against_torch_fp64(dev):
    """mcedm_edm_denoised / _euler_guided / _correct_guided against the torch expressions of ddim.py:1566-1592."""
    from mcedm_b200 import _lib as L
    from mcedm_b200.mcedm import precond_scalars

    lib = L.lib()
    gen = torch.Generator().manual_seed(9)
    B, C, H, W = 2, 1, 128, 128
    x_hat = torch.randn(B, C, H, W, generator=gen, dtype=torch.float64) * 3
    F1, F2 = torch.randn(B, C, H, W, generator=gen), torch.randn(B, C, H, W, generator=gen)
    dx1, dx2 = torch.randn(B, C, H, W, generator=gen) * 1e-4, torch.randn(B, C, H, W, generator=gen) * 1e-4
    t_cur, t_next = torch.tensor(2.5, dtype=torch.float64), torch.tensor(1.9, dtype=torch.float64)
    t_hat = t_cur + 0.3 * t_cur

    def D_of(xt, sig, Fx):
        cs, co, _, _ = O.precond_coeffs(sig)
        return cs * xt.float() + co * Fx

    d1 = D_of(x_hat, t_hat, F1)
    d_cur = (x_hat - d1.double()) / t_hat - 5. * dx1 / t_hat
    x_e = x_hat + (t_next - t_hat) * d_cur
    d2 = D_of(x_e, t_next, F2)
    d_prime = (x_e - d2.double()) / t_next - 5. * dx2 / t_hat
    x_new = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    xh, F1d, F2d, g1, g2 = [t.to(dev).contiguous() for t in (x_hat, F1, F2, dx1, dx2)]
    ones = torch.ones(B, C, H, W, device=dev)
    n = x_hat.numel()
    s = L.stream_ptr()
    cs, co, _, _ = precond_scalars(float(t_hat))
    cs2, co2, ci2, _ = precond_scalars(float(t_next))
    Db = torch.empty(B, C, H, W, device=dev)
    L.check(lib.mcedm_edm_denoised(L.ptr(xh), L.ptr(F1d), cs, co, n, L.ptr(Db), s))
    assert torch.equal(Db.cpu(), d1)
    dc, xe, xin = torch.empty_like(xh), torch.empty_like(xh), torch.empty_like(Db)
    L.check(lib.mcedm_edm_euler_guided(L.ptr(xh), L.ptr(Db), L.ptr(g1), L.ptr(ones), float(t_hat), float(t_next), ci2, n,
                                       L.ptr(dc), L.ptr(xe), L.ptr(xin), s))
    assert torch.equal(dc.cpu(), d_cur) and torch.equal(xe.cpu(), x_e)
    assert torch.equal(xin.cpu(), O.precond_coeffs(t_next)[2].reshape(()) * x_e.float())
    L.check(lib.mcedm_edm_denoised(L.ptr(xe), L.ptr(F2d), cs2, co2, n, L.ptr(Db), s))
    assert torch.equal(Db.cpu(), d2)
    xn = torch.empty_like(xh)
    L.check(lib.mcedm_edm_correct_guided(L.ptr(xh), L.ptr(xe), L.ptr(Db), L.ptr(g2), L.ptr(dc), L.ptr(ones),
                                         float(t_hat), float(t_next), n, L.ptr(xn), s))
    assert torch.equal(xn.cpu(), x_new)


def _cond_module(dev, g):
    from mcedm_b200.cond_edm import PlCondEdm
    from mcedm_b200.utils import randomize_zero_init

    cfg = hparams("config_adm_edm_res32_cond_h")
    torch.manual_seed(1)
    pl = PlCondEdm(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    sd = {k: v.detach().clone() for k, v in pl.model.state_dict().items()}
    pl = pl.to(dev)
    st = g["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    pl.set_pde_loss_function("swe_per", False)
    pl.eval()
    return pl, cfg, sd


def test_module_level_pde_loss_and_guidance(dev):
    """PlMcedm.get_pde_loss, PlCondEdm.get_pde_loss / get_dx_pde on the reference fixture; PlMcedm guide_dx raises."""
    from common import stress_module

    g = golden("pde.pt")
    gm = g["mcedm"]
    pl, cfg = stress_module()
    pl = pl.to(dev)
    st = gm["stats"]
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    pl.set_pde_loss_function("swe_per", False)
    pl.h_ch = pl.u_ch = 1
    h, u = D._FIELDS["swe_per"](2, 128, first_seed=gm["field_seed"])
    state = pl.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    gen = torch.Generator().manual_seed(gm["noise_seed"])
    sample = state.double() + 0.1 * torch.randn(state.shape, generator=gen, dtype=torch.float64).to(dev)
    for x, key in ((sample, "pde_sample"), (state, "pde_gt")):
        v = pl.get_pde_loss(x, clamp_loss=False, do_rearrange=False)
        assert v.dtype == torch.float32 and v.dim() == 0
        assert abs(float(v) - float(gm[key])) < 2e-6 * float(gm[key])
    v = pl.get_pde_loss(sample, do_rearrange=False)                   # clamp_loss=True default: matrix path
    assert abs(float(v) - float(gm["pde_sample_clamped"])) < 2e-6 * float(gm["pde_sample_clamped"])
    # `b c h w` input with do_rearrange=True is the same numbers
    v = pl.get_pde_loss(sample.permute(0, 3, 1, 2).contiguous(), clamp_loss=False)
    assert abs(float(v) - float(gm["pde_sample"])) < 2e-6 * float(gm["pde_sample"])
    with pytest.raises(NotImplementedError):                          # the reference raises RuntimeError here
        pl.sample_edm(sample.permute(0, 3, 1, 2).float(), state.permute(0, 3, 1, 2), torch.ones_like(state).permute(0, 3, 1, 2),
                      cfg.diff_sampler, guide_dx=True)

    gc = g["cond"]
    plc, cfgc, _ = _cond_module(dev, gc)
    h, u = D._FIELDS["swe_per"](1, 128, first_seed=gc["field_seed"])
    state = plc.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    h_n, u_n = state[..., :1], state[..., 1:2]
    gen = torch.Generator().manual_seed(gc["noise_seed"])
    u_s = u_n.double() + 0.1 * torch.randn(u_n.shape, generator=gen, dtype=torch.float64).to(dev)
    v = plc.get_pde_loss(h_n, u_s, clamp_loss=False, do_rearrange=False)
    assert abs(float(v) - float(gc["pde"])) < 2e-6 * float(gc["pde"])
    hc, uc = h_n.permute(0, 3, 1, 2), u_s.permute(0, 3, 1, 2)
    for calc_prob, key in ((True, "dx_mean"), (False, "dx_sum")):
        d = plc.get_dx_pde(hc, uc, calc_prob=calc_prob).cpu()
        assert d.shape == gc[key].shape
        assert float((d - gc[key]).abs().max()) < 2e-5 * float(gc[key].abs().max())
    assert torch.equal(plc.get_dx_log_prob(hc, uc, False), torch.zeros_like(uc))


def test_guided_cond_sampler_against_reference(dev):
    """3-step PDE-guided PlCondEdm.sample_edm: same RNG sequence, per-evaluation D_x within the 16-bit bar on the same
    input, and the final sample close to the reference's guided sample — closer than to its unguided one."""
    gc = golden("pde.pt")["cond"]
    pl, cfg, sd = _cond_module(dev, gc)
    mcfg = dict(cfg.model.hparams.model)
    h, u = D._FIELDS["swe_per"](1, 128, first_seed=gc["field_seed"])
    state = pl.data_transform(torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev))
    h_n, u_n = state[..., :1], state[..., 1:2]
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps = gc["sample"]["steps"]
    feed = NoiseFeed(gc["sample"]["seed"])
    pl._noise_hook = feed.hook
    pl._trace = []
    u_noise = feed.draw(u_n)
    xs = pl.sample_edm(pl.get_cond_in(h_n, u_n, None, None), u_noise, sp, return_last=True, guide_dx=True)
    assert [tuple(c) for c in feed.calls] == [tuple(c) for c in gc["sample"]["calls"]]
    assert xs.shape == (1, 1, 128, 128, 1) and xs.dtype == torch.float64
    cond_c = h_n.permute(0, 3, 1, 2).contiguous().cpu()
    assert len(pl._trace) == len(gc["sample"]["denoised"])
    for (i, which, sigma, d, xt), ref in zip(pl._trace, gc["sample"]["denoised"]):
        assert abs(sigma - ref["sigma"]) <= 1e-6 * max(1.0, ref["sigma"])
        with torch.no_grad():
            d_or, _ = O.denoise(sd, mcfg, xt.cpu(), torch.tensor(sigma, dtype=torch.float64), cond_c)
        assert rel_l2(d, d_or) < 1e-2
    e_g, e_p = rel_l2(xs, gc["sample"]["xs"]), rel_l2(xs, gc["sample"]["xs_plain"])
    assert e_g < 5e-2 and e_g < e_p
    # guidance off reproduces the unguided fixture the same way
    feed = NoiseFeed(gc["sample"]["seed"])
    pl._noise_hook, pl._trace = feed.hook, None
    u_noise = feed.draw(u_n)
    xs0 = pl.sample_edm(pl.get_cond_in(h_n, u_n, None, None), u_noise, sp, return_last=True, guide_dx=False)
    e_g, e_p = rel_l2(xs0, gc["sample"]["xs"]), rel_l2(xs0, gc["sample"]["xs_plain"])
    assert e_p < 5e-2 and e_p < e_g
    # test_step logs the residual metrics of the reference (ddim.py:1291-1303)
    pl._noise_hook = None
    sp.n_samples = 2
    pl.set_test_sampler_params(sp)
    h_t, u_t = torch.from_numpy(h).to(dev), torch.from_numpy(u).to(dev)
    out = pl.test_step((h_t, None, None, u_t), 0)
    for k in ("test_mae_u", "test_mae_u_un", "test_mae_u_scaled", "test_corr_u", "test_pde_loss", "test_pde_loss_gt"):
        assert k in pl.logged and torch.isfinite(torch.as_tensor(pl.logged[k])), k
    _, tot = P.get_pde_loss(state[..., 0].cpu().numpy(), state[..., 1].cpu().numpy(),
                            {k: float(v) for k, v in gc["stats"].items()}, "swe_per")
    assert abs(float(pl.logged["test_pde_loss_gt"]) - tot) < 2e-6 * tot
    assert out["traj"].shape == (1, 1, 128, 128, 2, 1)
