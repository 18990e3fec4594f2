"""sbi's `made` estimator (MADE with a mixture-of-Gaussians head, flow.py:37-112) on the NSF kernels with the
MoG head: the reference fixture (log_prob), sampling against the moments / quantiles of 20 000 reference
samples, gradients against the reference's fp64 autograd, and an NPE training run."""
import math
import os
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load():
    from sbi_b200.neural_nets import build_made
    g = torch.load(os.path.join(GOLD, "made_d3c2.pt"))
    est = build_made(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    return g, est.cuda()


def test_made_log_prob_reproduces_reference_fixture(cuda_lib):
    g, est = _load()
    with torch.no_grad():
        lp = est.log_prob(g["inp"].cuda(), g["cond"].cuda())[0].cpu()
        lps = est.log_prob(g["inp"].unsqueeze(1).cuda(), g["cond"][:1].cuda())[:, 0].cpu()
        big = est.log_prob(g["inp"].repeat(200, 1).unsqueeze(1).cuda(), g["cond"][:1].cuda())[:, 0].cpu()
    assert (lp - g["log_prob"]).abs().max() <= 2e-3
    assert (lps - g["log_prob_shared"]).abs().max() <= 2e-3
    assert (big.reshape(200, -1) - g["log_prob_shared"]).abs().max() <= 2e-3      # 12 800 rows: many tiles


def test_made_sampling_matches_reference_sample_statistics(cuda_lib):
    g, est = _load()
    torch.manual_seed(0)
    s = est.sample((100_000,), g["cond"][:1].cuda())[:, 0].cpu()
    assert s.shape == (100_000, 3) and torch.isfinite(s).all()
    se = g["sample_std"] / math.sqrt(20000)
    assert ((s.mean(0) - g["sample_mean"]).abs() < 6 * se + 1e-3).all()
    assert (s.std(0) / g["sample_std"] - 1).abs().max() < 0.05
    q = torch.quantile(s, torch.tensor([0.1, 0.5, 0.9]), dim=0)
    assert (q - g["sample_q"]).abs().max() < 0.06 * g["sample_std"].max()
    # the samples are draws of the density log_prob evaluates: E_q[log q] ~ entropy consistency with a second
    # batch (a wrong sampler shifts the average log-density)
    with torch.no_grad():
        a = est.log_prob(s[:20000].unsqueeze(1).cuda(), g["cond"][:1].cuda()).mean().item()
        b = est.log_prob(s[20000:40000].unsqueeze(1).cuda(), g["cond"][:1].cuda()).mean().item()
    assert abs(a - b) < 0.05


@pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")
def test_made_gradients_match_reference_autograd(cuda_lib):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    g, est = _load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_posterior_nn("made")(g["theta"], g["x"])
    ref.load_state_dict(g["state_dict"])
    ref = ref.double()
    R = 100
    inp, cond = g["theta"][:R] * 1.4, g["x"][:R]
    w = torch.randn(R, dtype=torch.float64)
    a, c = inp.double().requires_grad_(True), cond.double().requires_grad_(True)
    lp64 = ref.log_prob(a, c)[0]
    (lp64 * w).sum().backward()
    want = est.layout.pack({k: p.grad for k, p in ref.named_parameters()}).double()
    ai, ci = inp.cuda().requires_grad_(True), cond.cuda().requires_grad_(True)
    lp = est.log_prob(ai, ci)[0]
    (lp * w.float().cuda()).sum().backward()
    assert (lp.detach().cpu().double() - lp64.detach()).abs().max() <= 2e-3
    mask = est.net._mask.cpu().bool()
    got = est.flat.grad.cpu().double() * mask
    want = want * mask
    sc = want.abs().max()
    assert (got - want).abs().max() <= 2e-3 * sc, ((got - want).abs().max() / sc).item()
    assert (ai.grad.cpu().double() - a.grad).abs().max() <= 2e-3 * a.grad.abs().max()
    assert (ci.grad.cpu().double() - c.grad).abs().max() <= 2e-3 * c.grad.abs().max()


def test_npe_with_made_fits_linear_gaussian(cuda_lib):
    from torch.distributions import MultivariateNormal
    from sbi_b200.inference import NPE
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), torch.eye(D))
    theta = prior.sample((6000,))
    x = theta + math.sqrt(0.3) * torch.randn_like(theta)
    inf = NPE(prior, density_estimator="made", device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf.append_simulations(theta, x).train(training_batch_size=200, max_num_epochs=40)
    x_o = torch.tensor([[0.6, -0.4, 0.2]])
    s = inf.build_posterior().sample((4000,), x=x_o).cpu()
    assert (s.mean(0) - x_o[0] / 1.3).abs().max() < 0.08
    assert (s.std(0) / math.sqrt(0.3 / 1.3) - 1).abs().max() < 0.2
