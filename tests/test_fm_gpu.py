"""Flow-matching (FMPE) kernels against the CPU oracle, and an end-to-end FMPE fit."""
import math

import pytest
import torch

from oracle import sbi_port

pytestmark = pytest.mark.gpu


def _pair(D=5, C=7, seed=0, perturb=0.1, **kw):
    from sbi_b200.flowmatching import build_vector_field_estimator
    g = torch.Generator().manual_seed(seed)
    theta, x = 0.7 * torch.randn(500, D, generator=g) + 0.4, 1.5 * torch.randn(500, C, generator=g) - 0.3
    torch.manual_seed(seed)
    ref = sbi_port.build_flow_matching_estimator(theta, x, **kw)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(perturb * torch.randn(p.shape, generator=g))
    est = build_vector_field_estimator(theta, x, **{("num_layers" if k == "num_layers" else k): v for k, v in kw.items()})
    est.load_state_dict(ref.state_dict())
    return ref, est.cuda(), theta, x


@pytest.mark.parametrize("D,C,R", [(5, 7, 100), (20, 20, 1000), (2, 3, 17), (3, 2, 5000)])
def test_fm_forward_matches_oracle(cuda_lib, D, C, R):
    ref, est, theta, x = _pair(D, C)
    g = torch.Generator().manual_seed(2)
    inp = torch.randn(R, D, generator=g)
    cond = torch.randn(R, C, generator=g)
    t = torch.rand(R, generator=g)
    with torch.no_grad():
        v64 = ref.double().forward(inp.double(), cond.double(), t.double())
        v = est.forward(inp.cuda(), cond.cuda(), t.cuda()).cpu()
        # shared condition / shared time (the ODE right-hand side call pattern)
        v64s = ref.forward(inp.double(), cond[:1].double(), torch.tensor(0.3, dtype=torch.float64))
        vs = est.forward(inp.cuda(), cond[:1].cuda(), torch.tensor(0.3, device="cuda")).cpu()
    assert (v.double() - v64).abs().max() <= 2e-3 * max(1.0, v64.abs().max().item())
    assert (vs.double() - v64s).abs().max() <= 2e-3 * max(1.0, v64s.abs().max().item())


@pytest.mark.parametrize("D,C,R", [(5, 7, 64), (20, 20, 512), (2, 3, 17), (3, 2, 2000)])
def test_fm_loss_and_grads_match_oracle(cuda_lib, D, C, R):
    ref, est, theta, x = _pair(D, C)
    g = torch.Generator().manual_seed(3)
    inp, cond = theta[:R].clone(), x[:R].clone()
    if R > theta.shape[0]:
        inp, cond = torch.randn(R, D, generator=g), torch.randn(R, C, generator=g)
    t = torch.rand(R, generator=g)
    eps = torch.randn(R, D, generator=g)
    w = torch.randn(R, generator=g)

    def oracle(dtype):
        r = ref.to(dtype)
        r.zero_grad()
        l = r.loss(inp.to(dtype), cond.to(dtype), times=t.to(dtype), theta_1=eps.to(dtype))
        (l * w.to(dtype)).sum().backward()
        return l.detach().double(), est.layout.pack({k: p.grad for k, p in r.named_parameters() if k.startswith("net.")}).double()

    l32, g32 = oracle(torch.float32)
    l64, g64 = oracle(torch.float64)
    from sbi_b200.flowmatching import _FmLoss
    est.zero_grad()
    loss = _FmLoss.apply(est.net.flat, inp.cuda(), cond.cuda(), t.cuda(), eps.cuda(), est)
    (loss * w.cuda()).sum().backward()
    assert (loss.detach().cpu().double() - l64).abs().max() <= 1e-3 * max(1.0, l64.abs().max().item())
    sc = g64.abs().max().item()
    err, err32 = (est.flat.grad.cpu().double() - g64).abs().max().item() / sc, (g32 - g64).abs().max().item() / sc
    print(f"fm D={D} R={R}: grad rel err {err:.3e} (torch-fp32 {err32:.3e})")
    assert err <= max(2e-3, 4 * err32)


def test_dopri5_on_linear_ode():
    from sbi_b200.flowmatching import odeint_dopri5
    y0 = torch.tensor([[1.0, 2.0]], dtype=torch.float64)
    y, nfe = odeint_dopri5(lambda y, t: -y, y0, 0.0, 1.0)
    assert (y - y0 * math.exp(-1)).abs().max() < 1e-5 and nfe < 200
    y, _ = odeint_dopri5(lambda y, t: -y, y0, 1.0, 0.0)       # backwards in time
    assert (y - y0 * math.exp(1)).abs().max() < 1e-4


def test_fmpe_fit_linear_gaussian(cuda_lib):
    """tests/linearGaussian_vector_field_test.py:48-152 analogue (fmpe, gaussian prior, ODE sampling)."""
    from torch.distributions import MultivariateNormal
    from sbi_b200.inference import FMPE
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), 0.1 * torch.eye(D))
    theta = prior.sample((6000,))
    x = theta + math.sqrt(0.1) * torch.randn_like(theta)
    inf = FMPE(prior, device="cuda")
    inf.append_simulations(theta, x).train(training_batch_size=500, max_num_epochs=150)
    assert inf.summary["validation_loss"][-1] < inf.summary["validation_loss"][0]
    post = inf.build_posterior()
    x_o = torch.tensor([[0.3, -0.2, 0.1]])
    s = post.sample((3000,), x=x_o).cpu()
    assert (s.mean(0) - x_o[0] / 2).abs().max() < 0.05
    assert (s.std(0) / math.sqrt(0.05) - 1).abs().max() < 0.25
    # reverse-SDE sampling (Euler-Maruyama, 500 steps) of the same estimator:
    # linearGaussian_vector_field_test.py samples fmpe with both "ode" and "sde"
    post_sde = inf.build_posterior(sample_with="sde")
    s2 = post_sde.sample((3000,), x=x_o).cpu()
    assert torch.isfinite(s2).all()
    assert (s2.mean(0) - x_o[0] / 2).abs().max() < 0.06
    assert (s2.std(0) / math.sqrt(0.05) - 1).abs().max() < 0.3


def test_fm_score_drift_diffusion_formulas(cuda_lib):
    """score / drift / diffusion / schedule of the estimator (flowmatching_estimator.py:374-469,
    estimators/base.py:605-624) against the oracle port's velocity on the same weights."""
    ref, est, theta, x = _pair(4, 3)
    g = torch.Generator().manual_seed(3)
    th = torch.randn(64, 4, generator=g)
    for tv in (0.05, 0.5, 0.93):          # the predictor passes a 0-dim time (ts[i])
        t = torch.tensor(tv)
        with torch.no_grad():
            v = ref.double()(th.double(), x[:64].double(), t.double())
            score_ref = (-(1 - tv) * v - th.double()) / (tv + 1e-3)
        score_k = est.score(th.cuda(), x[:64].cuda(), t.cuda())
        assert score_k.shape == (64, 4)
        assert (score_k.cpu().double() - score_ref).abs().max() <= 5e-3 * score_ref.abs().max()
    f = est.drift_fn(th.cuda(), torch.tensor(0.995, device="cuda"))
    assert torch.allclose(f.cpu(), -th / 0.01, rtol=1e-4)
    gdiff = est.diffusion_fn(th.cuda(), torch.tensor(0.5, device="cuda"))
    assert abs(float(gdiff) - math.sqrt(2 * (0.5 + 1e-3) / 0.5)) < 1e-5
    ts = est.solve_schedule(11)
    assert float(ts[0]) == 1.0 and float(ts[-1]) == 0.0 and ts.numel() == 11
