"""The ONE-dimensional neural spline flow (scalar input; conditioner = context-only MLP, flow.py:401-432 and
ContextSplineMap :1419-1478) on the NSF kernels (cond_mlp = 1): the reference fixture (log_prob, shared
condition, inverse_transform, sampling inverse + log|det|), gradients against the reference's fp64 autograd,
and NLE-style training of a scalar likelihood."""
import math
import os
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load():
    from sbi_b200.neural_nets import build_nsf
    g = torch.load(os.path.join(GOLD, "nsf_d1c3.pt"))
    est = build_nsf(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    return g, est.cuda()


def test_nsf_1d_kernels_reproduce_reference_fixture(cuda_lib):
    g, est = _load()
    inp, cond, noise = g["inp"].cuda(), g["cond"].cuda(), g["noise"].cuda()
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0].cpu()
        lps = est.log_prob(inp.unsqueeze(1), cond[:1])[:, 0].cpu()
        z = est.inverse_transform(inp, cond).cpu()
        s, lad = est.inverse_flow(noise, cond)
        big = est.log_prob(inp.repeat(300, 1).unsqueeze(1), cond[:1])[:, 0].cpu()     # 19 200 rows
    assert (lp - g["log_prob"]).abs().max() <= 2e-3
    assert (lps - g["log_prob_shared"]).abs().max() <= 2e-3
    assert (z - g["inverse_transform"]).abs().max() <= 2e-3
    assert (s.cpu() - g["samples"]).abs().max() <= 2e-3
    assert (lad.cpu() - g["inverse_logabsdet"]).abs().max() <= 5e-3
    assert (big.reshape(300, -1) - g["log_prob_shared"]).abs().max() <= 2e-3


@pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")
@pytest.mark.parametrize("hl", [1, 2])
def test_nsf_1d_gradients_match_reference_autograd(cuda_lib, hl):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    from sbi_b200.neural_nets import posterior_nn
    gen = torch.Generator().manual_seed(2)
    theta = 0.7 * torch.randn(400, 1, generator=gen) + 0.3
    x = 1.3 * torch.randn(400, 4, generator=gen) - 0.2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(3)
        ref = ref_posterior_nn("nsf", hidden_layers_spline_context=hl)(theta, x)
        est = posterior_nn("nsf", hidden_layers_spline_context=hl)(theta, x)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=gen))
    est.load_state_dict(ref.state_dict())
    est = est.cuda()
    ref = ref.double()
    R = 150
    inp, cond = theta[:R] * 1.6, x[:R]
    w = torch.randn(R, dtype=torch.float64)
    a, c = inp.double().requires_grad_(True), cond.double().requires_grad_(True)
    lp64 = ref.log_prob(a, c)[0]
    (lp64 * w).sum().backward()
    # weight-shared layers: the reference holds ONE parameter under several names; take each tensor once
    grads = {}
    for k, p in ref.named_parameters():
        grads[k] = p.grad
    sd_keys = [k for k in est.layout.index]
    named = dict(ref.named_parameters())
    full = {}
    for k in sd_keys:
        kk = k
        if kk not in named:      # alias of the shared hidden layer -> its first registration
            base = kk.rsplit("spline_predictor.", 1)
            kk = base[0] + "spline_predictor.2." + kk.rsplit(".", 1)[1]
        full[k] = named[kk].grad
    want = est.layout.pack(full).double()
    ai, ci = inp.cuda().requires_grad_(True), cond.cuda().requires_grad_(True)
    lp = est.log_prob(ai, ci)[0]
    (lp * w.float().cuda()).sum().backward()
    assert (lp.detach().cpu().double() - lp64.detach()).abs().max() <= 2e-3
    got = est.flat.grad.cpu().double()
    sc = want.abs().max()
    assert (got - want).abs().max() <= 2e-3 * sc, ((got - want).abs().max() / sc).item()
    assert (ai.grad.cpu().double() - a.grad).abs().max() <= 2e-3 * a.grad.abs().max()
    assert (ci.grad.cpu().double() - c.grad).abs().max() <= 2e-3 * c.grad.abs().max()


def test_scalar_density_fit(cuda_lib):
    """NPE with scalar theta: theta | x ~ N(., .) analytically; the trained 1-D flow recovers location and
    spread (median / interquartile range: a 10-bin spline leaves light artefacts in the far tails)."""
    from torch.distributions import MultivariateNormal
    from sbi_b200.inference import NPE
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(1), 4.0 * torch.eye(1))
    theta = prior.sample((20000,))
    x = theta + 1.5 * torch.randn(20000, 3)
    inf = NPE(prior, density_estimator="nsf", device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf.append_simulations(theta, x).train(training_batch_size=500, max_num_epochs=80)
    vl = inf.summary["validation_loss"]
    assert vl[-1] < vl[0] - 0.3
    x_o = torch.tensor([[0.8, 1.0, 1.2]])
    s = inf.build_posterior().sample((20000,), x=x_o).cpu().reshape(-1)
    post_var = 1 / (1 / 4.0 + 3 / 2.25)
    post_mean = post_var * (3.0 / 2.25)
    q = torch.quantile(s, torch.tensor([0.25, 0.5, 0.75]))
    assert abs(q[1].item() - post_mean) < 0.12
    assert abs((q[2] - q[0]).item() / 1.349 / math.sqrt(post_var) - 1) < 0.25
    # and the density it evaluates is the density it samples: average log q of fresh samples ~ -entropy
    with torch.no_grad():
        lq = inf.build_posterior().log_prob(s[:5000].reshape(-1, 1), x=x_o).mean().item()
    assert abs(lq + 0.5 * math.log(2 * math.pi * math.e * post_var)) < 0.25
