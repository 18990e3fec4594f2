"""MAF with rational-quadratic-spline element-wise maps (`maf_rqs`, flow.py:212-330) on the CUDA MADE
kernels with the spline head: the reference fixture (log_prob, sampling inverse, log|det|), gradients
against the reference's fp64 autograd (reference imported through oracle.ref_shim when present), the
inverse o forward round trip at 2^17 rows, and a short NPE training run."""
import math
import os
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load():
    from sbi_b200.neural_nets import build_maf_rqs
    g = torch.load(os.path.join(GOLD, "maf_rqs_d4c3.pt"))
    est = build_maf_rqs(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    return g, est.cuda()


def test_maf_rqs_kernels_reproduce_reference_fixture(cuda_lib):
    g, est = _load()
    with torch.no_grad():
        lp = est.log_prob(g["inp"].cuda(), g["cond"].cuda())[0].cpu()
        s, lad = est.inverse_flow(g["noise"].cuda(), g["cond"].cuda())
    assert (lp - g["log_prob"]).abs().max() <= 2e-3
    assert (s.cpu() - g["samples"]).abs().max() <= 2e-3
    assert (lad.cpu() - g["inverse_logabsdet"]).abs().max() <= 5e-3


def test_maf_rqs_round_trip_and_bulk_rows(cuda_lib):
    """z = T(x | c) then T^{-1}(z | c) = x for 2^17 rows spanning the tails (|x| > tail bound)."""
    g, est = _load()
    R = 1 << 17
    gen = torch.Generator().manual_seed(3)
    inp = (2.5 * torch.randn(R, 4, generator=gen)).cuda()
    cond = g["x"][torch.randint(0, 400, (R,), generator=gen)].cuda()
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0]
        z = est.inverse_transform(inp, cond)                    # noise of the rows
        # inverse_transform standardises the condition itself only when told to: use the flow pair
        lp2, noise = est._logprob_raw(inp, cond, False, want_noise=True)
        back, lad = est.inverse_flow(noise, cond)
    assert torch.isfinite(lp).all() and torch.isfinite(z).all()
    assert (back - inp).abs().max() <= 2e-3
    base = -0.5 * (noise ** 2).sum(1) - 0.5 * 4 * math.log(2 * math.pi)
    assert (lp2 - (base - lad)).abs().max() <= 5e-3           # log q = log N(z) + log|dz/dx| = base - log|dx/dz|


@pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")
def test_maf_rqs_gradients_match_reference_autograd(cuda_lib):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    g, est = _load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_posterior_nn("maf_rqs")(g["theta"], g["x"])
    ref.load_state_dict(g["state_dict"])
    ref = ref.double()
    R = 200
    inp, cond = (g["theta"][:R] * 1.4), g["x"][:R]
    w = torch.randn(R, dtype=torch.float64)
    a = inp.double().requires_grad_(True)
    c = cond.double().requires_grad_(True)
    lp64 = ref.log_prob(a, c)[0]
    (lp64 * w).sum().backward()
    want = est.layout.pack({k: p.grad for k, p in ref.named_parameters()}).double()
    ai, ci = inp.cuda().requires_grad_(True), cond.cuda().requires_grad_(True)
    lp = est.log_prob(ai, ci)[0]
    (lp * w.float().cuda()).sum().backward()
    assert (lp.detach().cpu().double() - lp64.detach()).abs().max() <= 2e-3
    # the kernels compute dense weight gradients; masked-out entries are frozen by the Adam mask and must be
    # ignored here exactly like nflows' `weight * mask` zeroes them
    mask = est.net._mask.cpu().bool()
    got = est.flat.grad.cpu().double() * mask
    want = want * mask
    sc = want.abs().max()
    assert (got - want).abs().max() <= 2e-3 * sc, ((got - want).abs().max() / sc).item()
    assert (ai.grad.cpu().double() - a.grad).abs().max() <= 2e-3 * a.grad.abs().max()
    assert (ci.grad.cpu().double() - c.grad).abs().max() <= 2e-3 * c.grad.abs().max()


def test_npe_with_maf_rqs_fits_linear_gaussian(cuda_lib):
    from torch.distributions import MultivariateNormal
    from sbi_b200.inference import NPE
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), torch.eye(D))
    theta = prior.sample((6000,))
    x = theta + math.sqrt(0.3) * torch.randn_like(theta)
    inf = NPE(prior, density_estimator="maf_rqs", device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf.append_simulations(theta, x).train(training_batch_size=200, max_num_epochs=40)
    x_o = torch.tensor([[0.6, -0.4, 0.2]])
    s = inf.build_posterior().sample((4000,), x=x_o).cpu()
    assert (s.mean(0) - x_o[0] / 1.3).abs().max() < 0.08
    assert (s.std(0) / math.sqrt(0.3 / 1.3) - 1).abs().max() < 0.2
