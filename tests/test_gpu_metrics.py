"""GPU parity of the fused validation / test reductions (csrc/metrics.cu) against the torch expressions of
the reference (models/losses.py:62-78 MaskedLoss, :96-128 CorrelationLoss; models/mcedm.py:385-408; models/ddim.py:689-698):
fp64 arithmetic in the same per-element order, so 1e-12 relative for the masked errors, bit-exact min / max, 1e-6 for the
correlation (the reference sums the fp32 target's variance in fp32)."""
import copy

import pytest
import torch
from einops import rearrange

from common import stress_module
from mcedm_b200 import data as D
from mcedm_b200.nn_misc import CorrelationLoss, MaskedLoss, Normalizer, fused_masked_mae

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _ref_losses(xs_last, n, gt, h_un, u_un, mask, loss_dim, nh, nu, clamp=False):
    xs_mean = torch.mean(rearrange(xs_last, "(n b) h w c -> n b h w c", n=n), dim=0)
    crit = MaskedLoss()
    l1 = crit(xs_mean, gt, mask, loss_dim)
    h, u = xs_mean[..., 0:1], xs_mean[..., 1:2]
    if clamp:
        h, u = torch.clamp(h, 0.0, 1.0), torch.clamp(u, 0.0, 1.0)
    un = torch.cat([nh(h, inverse=True), nu(u, inverse=True)], dim=-1)
    l2 = crit(un, torch.cat([h_un, u_un], dim=-1), mask, loss_dim)
    return l1, l2


@pytest.mark.parametrize("n,b,H,W,c0,c1,clamp", [(1, 3, 128, 128, 0, 2, False), (4, 2, 128, 128, 1, 2, False),
                                                  (3, 2, 32, 16, 0, 1, True), (2, 1, 8, 8, 0, 2, False)])
def test_masked_mae_mean_matches_torch(dev, n, b, H, W, c0, c1, clamp):
    g = torch.Generator().manual_seed(n * 7 + b)
    xs = torch.randn(n * b, H, W, 2, generator=g, dtype=torch.float64).to(dev)
    gt = torch.randn(b, H, W, 2, generator=g).to(dev)
    h_un, u_un = torch.randn(b, H, W, 1, generator=g).to(dev) * 0.3 + 1.5, torch.randn(b, H, W, 1, generator=g).to(dev)
    mask = (torch.rand(b, H, W, 2, generator=g) > 0.4).float().to(dev)
    nh = Normalizer(torch.tensor(1.5), torch.tensor(0.3)).to(dev)
    nu = Normalizer(torch.tensor(0.1), torch.tensor(0.2)).to(dev)
    loss_dim = torch.arange(c0, c1).long() if (c0, c1) != (0, 2) else None
    got = fused_masked_mae(xs, n, gt, h_un, u_un, mask, c0, c1, nh, nu, clamp01=clamp)
    assert got is not None
    r1, r2 = _ref_losses(xs, n, gt, h_un, u_un, mask, loss_dim, nh, nu, clamp)
    assert got[0].dtype == torch.float64 and r1.dtype == torch.float64
    assert abs(float(got[0]) - float(r1)) <= 1e-12 * abs(float(r1)), (float(got[0]), float(r1))
    assert abs(float(got[1]) - float(r2)) <= 1e-12 * abs(float(r2)), (float(got[1]), float(r2))
    # an all-zero mask divides by zero exactly as torch does (nan), and other dtypes fall back (None)
    z = fused_masked_mae(xs, n, gt, h_un, u_un, torch.zeros_like(mask), c0, c1, nh, nu)
    assert torch.isnan(z[0])
    assert fused_masked_mae(xs.float(), n, gt, h_un, u_un, mask, c0, c1, nh, nu) is None


def test_correlation_and_minmax_match_torch(dev):
    from mcedm_b200.cond_edm import PlCondEdm

    g = torch.Generator().manual_seed(5)
    pred = (torch.randn(3, 128, 128, 2, generator=g, dtype=torch.float64) * 2 + 0.3).to(dev)
    tgt = (0.7 * pred.float().cpu() + 0.5 * torch.randn(3, 128, 128, 2, generator=g)).to(dev)
    tgt[2, :, :, 1] = 0.25                                            # constant target channel: zero denominator branch
    corr = CorrelationLoss()(pred, tgt)
    p, t = pred.reshape(3, -1, 2), tgt.reshape(3, -1, 2)
    ref = CorrelationLoss.calculate_correlation(p, t)
    assert corr.dtype == torch.float64 and corr.shape == (2,)
    assert torch.allclose(corr, ref.double(), rtol=1e-6, atol=1e-9), (corr, ref)
    scaled, lo, hi = PlCondEdm.scale_each_min_max(pred, return_min_max=True)
    flat = rearrange(pred, "b h w c -> b c (h w)")
    assert torch.equal(lo, flat.min(dim=2, keepdim=True)[0]) and torch.equal(hi, flat.max(dim=2, keepdim=True)[0])
    assert torch.equal(scaled, rearrange((flat - lo) / (hi - lo), "b c (h w) -> b h w c", h=128, w=128))


def test_test_step_metrics_fused_equal_torch_path(dev):
    """PlMcedm.test_step / validation_step log the same masked errors through the kernel and through the torch expressions
    (same sampled fields: the noise is injected)."""
    from common import NoiseFeed

    pl, cfg = stress_module()
    pl = pl.to(dev).eval()
    st = D.field_stats("swe_per", 16)
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    sp = copy.deepcopy(cfg.diff_sampler)
    sp.timesteps, sp.n_samples = 2, 3
    pl.set_test_sampler_params(sp)
    pl.sparams = sp
    batch = tuple(t.to(dev) if torch.is_tensor(t) else {k: v.to(dev) for k, v in t.items()}
                  for t in D.make_batch("swe_per", 2, "eval", seed=3))
    logs = []
    for fused in (True, False):
        pl.fused_metrics = fused
        pl._noise_hook = NoiseFeed(11).hook
        with torch.no_grad():
            out = pl.test_step(batch, 0)
            pl._noise_hook = NoiseFeed(12).hook
            outv = pl.validation_step(batch, 0)
        logs.append({k: float(v) for k, v in pl.logged.items() if "mae" in k})
        assert all(torch.isfinite(torch.as_tensor(v)) for v in logs[-1].values())
        assert out["loss_u"].dtype == torch.float64 and outv["loss_u_un"].dtype == torch.float64
    assert set(logs[0]) == set(logs[1]) and len(logs[0]) >= 8
    for k in logs[0]:
        assert abs(logs[0][k] - logs[1][k]) <= 1e-12 * abs(logs[1][k]), (k, logs[0][k], logs[1][k])
