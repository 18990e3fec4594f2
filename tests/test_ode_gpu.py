"""ODE path of the flow-matching posterior on the device (csrc/ode.cu step control, csrc/fm.cu exact-trace
kernel): the divergence kernel against autograd on the oracle's estimator, the device-controlled Dormand-
Prince against an analytic ODE and against the host-loop solver, and the neural-ODE log-probability
against the oracle's exact-trace CNF (oracle/ode_port.py) at atol 1e-6 / rtol 1e-5."""
import math

import pytest
import torch

from oracle import ode_port
from tests.test_fm_gpu import _pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,C,R", [(5, 7, 100), (20, 20, 333), (2, 3, 17)])
def test_velocity_divergence_matches_autograd(cuda_lib, D, C, R):
    ref, est, theta, x = _pair(D, C)
    ref = ref.double()
    g = torch.Generator().manual_seed(4)
    inp = torch.randn(R, D, generator=g)
    cond = torch.randn(1, C, generator=g)
    for t in (0.0, 0.37, 1.0):
        tt = torch.tensor(t, dtype=torch.float64)
        v64, div64 = ode_port.divergence_exact(lambda y, _t: ref.forward(y, cond.double(), tt), inp.double(), t)
        v, div = est.forward_and_divergence(inp.cuda(), cond.cuda(), torch.tensor(t, device="cuda"))
        assert (v.cpu().double() - v64).abs().max() <= 2e-3 * max(1.0, v64.abs().max().item())
        assert (div.cpu().double() - div64).abs().max() <= 2e-3 * max(1.0, div64.abs().max().item()), t
    # per-row times and per-row conditions
    tr = torch.rand(R, generator=g)
    cr = torch.randn(R, C, generator=g)
    v, div = est.forward_and_divergence(inp.cuda(), cr.cuda(), tr.cuda())
    with torch.enable_grad():
        yy = inp.double().requires_grad_(True)
        vv = ref.forward(yy, cr.double(), tr.double())
        d64 = sum(torch.autograd.grad(vv[:, i].sum(), yy, retain_graph=True)[0][:, i] for i in range(D))
    assert (div.cpu().double() - d64).abs().max() <= 2e-3 * max(1.0, d64.abs().max().item())


@pytest.mark.parametrize("use_graph", [False, True])
def test_device_dopri5_on_analytic_ode(cuda_lib, use_graph):
    """y' = -1.5 y + sin t with a torch right-hand side reading the stage time from the device scalar."""
    from sbi_b200.flowmatching import DeviceDopri5
    n = 5000
    y0 = torch.linspace(-2, 2, n, device="cuda")
    tbuf = {}

    def rhs(y, t_ptr, out):
        t = tbuf["t"]                                 # view of the control block's t_stage entry
        torch.add(-1.5 * y, torch.sin(t), out=out)

    solver = DeviceDopri5(n, "cuda", rhs, atol=1e-7, rtol=1e-6, use_graph=use_graph)
    tbuf["t"] = solver.ctrl[6:7]
    y, nfe, steps, acc = solver.solve(y0, 0.0, 2.0)
    a, t = 1.5, 2.0
    want = (y0 + 1 / (a * a + 1)) * math.exp(-a * t) + (a * math.sin(t) - math.cos(t)) / (a * a + 1)
    assert (y - want).abs().max() < 2e-5
    assert nfe == 1 + 6 * steps and acc <= steps and steps < 200
    yb, *_ = solver.solve(y.clone(), 2.0, 0.0)        # backwards with the same (captured) step graph
    assert (yb - y0).abs().max() < 2e-4


def test_device_control_equals_host_loop_for_sampling(cuda_lib):
    from sbi_b200.flowmatching import sample_ode
    ref, est, theta, x = _pair(5, 7)
    torch.manual_seed(3)
    a, nfe_a = sample_ode(est, 2000, x[:1], return_nfe=True, device_control=True)
    torch.manual_seed(3)
    b, nfe_b = sample_ode(est, 2000, x[:1], return_nfe=True, device_control=False)
    assert torch.isfinite(a).all()
    assert (a - b).abs().max() < 1e-3             # same algorithm; fp32 reduction order of the error norm differs
    assert abs(nfe_a - nfe_b) <= 12


def test_neural_ode_log_prob_matches_oracle_cnf(cuda_lib):
    from sbi_b200.flowmatching import log_prob_ode
    D, C, R = 4, 3, 64
    ref, est, theta, x = _pair(D, C)
    ref = ref.double()
    g = torch.Generator().manual_seed(5)
    th = 0.8 * torch.randn(R, D, generator=g) + 0.3
    cond = x[:1]
    want, nfe64 = ode_port.cnf_log_prob(
        lambda y, t: ref.forward(y, cond.double(), torch.tensor(t, dtype=torch.float64)), th.double(), 0.0, 1.0,
        torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64), atol=1e-6, rtol=1e-5)
    got, nfe = log_prob_ode(est, th.cuda(), cond.cuda(), atol=1e-6, rtol=1e-5, return_nfe=True)
    print(f"neural-ODE log_prob: max |d| {(got.cpu().double() - want).abs().max():.2e}, nfe {nfe} (oracle {nfe64})")
    assert (got.cpu().double() - want).abs().max() < 5e-3


def test_vector_field_posterior_log_prob_linear_gaussian(cuda_lib):
    """FMPE on the linear-Gaussian task: posterior.log_prob (neural ODE, exact trace) is close to the
    analytic posterior density and -inf outside the prior support."""
    import warnings
    from torch.distributions import Independent, MultivariateNormal, Uniform
    from sbi_b200.inference import FMPE
    D = 2
    torch.manual_seed(0)
    prior = Independent(Uniform(-3 * torch.ones(D), 3 * torch.ones(D)), 1)
    theta = prior.sample((20_000,))
    x = theta + math.sqrt(0.1) * torch.randn_like(theta)
    inf = FMPE(prior, device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf.append_simulations(theta, x).train(training_batch_size=1000, max_num_epochs=60)
    post = inf.build_posterior()
    x_o = torch.tensor([[0.5, -0.3]])
    true = MultivariateNormal(x_o[0], 0.1 * torch.eye(D))
    th = true.sample((500,))
    lp = post.log_prob(th, x=x_o).cpu()
    assert torch.isfinite(lp).all()
    assert (lp - true.log_prob(th)).abs().mean() < 0.25
    assert post.log_prob(torch.tensor([[3.5, 0.0]]), x=x_o).item() == float("-inf")


def test_sde_step_kernel_matches_reference_formula(cuda_lib):
    """csrc/ode.cu `sde_em_step_kernel` against the reference's Euler-Maruyama arithmetic in torch ops
    (predictors.py:112-120 on flowmatching_estimator.py:374-469), same velocity, same normal draw, for
    three consecutive grid points (the control block advances on the device)."""
    import ctypes as C
    from sbi_b200 import _lib as L
    lib = L.load()
    ref, est, theta, x = _pair(5, 7)
    g = torch.Generator().manual_seed(6)
    th = torch.randn(300, 5, generator=g).cuda()
    cond = x[:1].cuda()
    ts = torch.linspace(1.0, 0.0, 9).cuda()
    eta = 0.8
    got = th.clone().contiguous()
    want = th.clone()
    ctrl = torch.tensor([1.0, 1.0], device="cuda")
    for i in range(1, 4):
        z = torch.randn(300, 5, generator=g).cuda()
        t1, t0 = ts[i - 1], ts[i]
        dt = t1 - t0
        f = est.drift_fn(want, t1)
        gg = est.diffusion_fn(want, t1)
        score = est.score(want, cond, t1)
        want = want - (f - (1 + eta ** 2) / 2 * gg ** 2 * score) * dt + (eta * gg) * z * torch.sqrt(dt)
        v = est.forward(got, cond, ctrl[0:1].clone())
        L.check(lib.sbi_b200_sde_em_step(got.data_ptr(), v.contiguous().data_ptr(), z.data_ptr(), got.numel(),
                                         ts.data_ptr(), ctrl.data_ptr(), eta, float(est.noise_scale), 0.99,
                                         L.stream_ptr()), "sde_em_step")
        assert abs(ctrl[0].item() - t0.item()) < 1e-7 and int(ctrl[1].item()) == i + 1
    assert (got - want).abs().max() <= 1e-4 * max(1.0, want.abs().max().item())


def test_fused_sde_sampler_matches_eager_in_distribution(cuda_lib):
    from sbi_b200.flowmatching import sample_sde
    ref, est, theta, x = _pair(3, 2, perturb=0.02)
    torch.manual_seed(0)
    a = sample_sde(est, 20000, x[:1], steps=100, fused=True)
    b = sample_sde(est, 20000, x[:1], steps=100, fused=False)
    assert torch.isfinite(a).all()
    assert (a.mean(0) - b.mean(0)).abs().max() < 0.05 * max(1.0, b.std(0).max().item())
    assert (a.std(0) / b.std(0) - 1).abs().max() < 0.05
