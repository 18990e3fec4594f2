"""Pins of oracle/ode_port.py (the restated zuko dopri5 + exact-trace CNF): an ODE with a closed-form
solution and the closed-form log-density of an affine (Gaussian) flow."""
import math

import torch

from oracle import ode_port


def test_dopri5_solves_linear_ode_to_tolerance():
    y0 = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    y, nfe = ode_port.odeint_dopri5(lambda y, t: -1.5 * y + math.sin(t), y0, 0.0, 2.0, atol=1e-9, rtol=1e-8)
    # y' = -a y + sin t: y = (y0 - c) e^{-a t} + (a sin t - cos t) / (a^2 + 1), c = -1 / (a^2 + 1)
    a, t = 1.5, 2.0
    want = (y0 + 1 / (a * a + 1)) * math.exp(-a * t) + (a * math.sin(t) - math.cos(t)) / (a * a + 1)
    assert (y - want).abs().max() < 1e-7 and nfe < 400
    back, _ = ode_port.odeint_dopri5(lambda y, t: -1.5 * y + math.sin(t), y, 2.0, 0.0, atol=1e-9, rtol=1e-8)
    assert (back - y0).abs().max() < 1e-6


def test_cnf_log_prob_of_affine_flow_is_gaussian():
    """v(y, t) = A y + b (constant in t): z = e^{A} theta + ..., log|det| = trace(A); the density of theta is
    the Gaussian pull-back of the base."""
    torch.manual_seed(0)
    D = 3
    A = 0.3 * torch.randn(D, D, dtype=torch.float64)
    b = torch.randn(D, dtype=torch.float64)
    theta = torch.randn(7, D, dtype=torch.float64)
    mu, sd = torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64)
    lp, _ = ode_port.cnf_log_prob(lambda y, t: y @ A.T + b, theta, 0.0, 1.0, mu, sd, atol=1e-10, rtol=1e-9)
    E = torch.linalg.matrix_exp(A)
    # z = E theta + (int_0^1 e^{A s} ds) b
    n = 2000
    s = (torch.arange(n, dtype=torch.float64) + 0.5) / n
    integ = sum(torch.linalg.matrix_exp(A * si) for si in s) / n
    z = theta @ E.T + integ @ b
    want = (-0.5 * z ** 2 - 0.5 * math.log(2 * math.pi)).sum(1) + torch.trace(A)
    assert (lp - want).abs().max() < 1e-6


def test_tableau_and_solution_agree_with_scipy_rk45():
    """scipy's RK45 is the same Dormand-Prince 5(4) pair (an independent third-party implementation in this
    image): identical Butcher tableau and error weights, and the same solution of a nonlinear system to within the
    tolerances (the step-size controllers differ in their error norm, so steps are not identical)."""
    import numpy as np
    from scipy.integrate import solve_ivp
    from scipy.integrate._ivp.rk import RK45
    for i in range(1, 6):
        assert np.allclose(RK45.A[i][:i], ode_port._A[i], rtol=0, atol=1e-16)
    assert np.allclose(RK45.A[1:6, :5][4], ode_port._A[5], atol=1e-16)
    assert np.allclose(RK45.C, ode_port._C[:6], atol=1e-16)
    assert np.allclose(RK45.B, ode_port._B5[:6], atol=1e-16) and np.allclose(ode_port._A[6], ode_port._B5[:6])
    e_port = np.array(ode_port._B5) - np.array(ode_port._B4)
    assert np.allclose(RK45.E, e_port, atol=1e-16) or np.allclose(RK45.E, -e_port, atol=1e-16)

    def rhs(t, y):           # Lotka-Volterra with forcing
        return np.array([1.1 * y[0] - 0.4 * y[0] * y[1] + 0.1 * np.sin(t), 0.1 * y[0] * y[1] - 0.4 * y[1]])

    y0 = np.array([10.0, 5.0])
    sp = solve_ivp(rhs, (0.0, 8.0), y0, method="RK45", rtol=1e-10, atol=1e-12)
    f = lambda y, t: torch.stack([1.1 * y[0] - 0.4 * y[0] * y[1] + 0.1 * math.sin(t), 0.1 * y[0] * y[1] - 0.4 * y[1]])
    y, nfe = ode_port.odeint_dopri5(f, torch.tensor(y0, dtype=torch.float64), 0.0, 8.0, atol=1e-12, rtol=1e-10)
    assert np.allclose(y.numpy(), sp.y[:, -1], rtol=1e-7, atol=1e-8), (y.numpy(), sp.y[:, -1])
