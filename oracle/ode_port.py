"""oracle/ode_port -- TEST INFRASTRUCTURE ONLY.

CPU restatement (fp64-capable, plain torch) of the ODE solver and the exact-trace continuous normalising
flow the reference delegates to the third-party package ``zuko==1.6.0`` (absent from /root/reference and
not installable offline; pinned in /root/reference/uv.lock): ``zuko.utils.odeint`` (adaptive
Dormand-Prince 5(4)) and ``zuko.transforms.FreeFormJacobianTransform(exact=True)`` wrapped in
``zuko.distributions.NormalizingFlow(transform, DiagNormal)``; reference call sites
/root/reference/sbi/samplers/ode_solvers/zuko_ode.py:80-124 and
/root/reference/sbi/inference/potentials/vector_field_potential.py:145-212.

Published algorithm restated: Dormand & Prince (1980) tableau with first-same-as-last reuse; one step size
for the whole (packed) state; error ratio = RMS over all entries of err / (atol + rtol * max(|y|, |y_new|));
accept iff <= 1; next step = h * clip(0.9 * ratio^(-1/5), 0.2, 5); log-density of the flow =
base.log_prob(y(t1)) + integral of trace(d f / d y), the trace taken exactly with autograd (one backward
pass per dimension).  **Parity unpinned** against the real zuko (no golden vectors in the reference);
pinned here by an analytic ODE, by the analytic log-density of an affine vector field and by scipy's RK45
(same tableau and error weights, same solution of a nonlinear system; tests/test_ode_port_cpu.py).
"""
import math

import torch

_A = [[], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9],
      [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
      [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
      [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_B5 = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_B4 = [5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40]


def odeint_dopri5(f, y0, t0, t1, atol=1e-6, rtol=1e-5, max_steps=100_000, h0=0.05):
    """y(t1) of y' = f(y, t), y(t0) = y0 (one flat tensor); returns (y, number of f evaluations)."""
    direction = 1.0 if t1 >= t0 else -1.0
    t, y = float(t0), y0
    h = direction * min(abs(t1 - t0), h0)
    k1 = f(y, t)
    nfe = 1
    for _ in range(max_steps):
        if (t1 - t) * direction <= 1e-12:
            break
        if (t + h - t1) * direction > 0:
            h = t1 - t
        ks = [k1]
        for i in range(1, 7):
            yi = y + h * sum(a * k for a, k in zip(_A[i], ks))
            ks.append(f(yi, t + _C[i] * h))
        nfe += 6
        y5 = y + h * sum(b * k for b, k in zip(_B5, ks))
        err = h * sum((b5 - b4) * k for b5, b4, k in zip(_B5, _B4, ks))
        tol = atol + rtol * torch.maximum(y.abs(), y5.abs())
        en = float(torch.sqrt(torch.mean((err / tol) ** 2)))
        if en <= 1.0:
            t, y, k1 = t + h, y5, ks[6]
        fac = 0.9 * (1.0 / max(en, 1e-10)) ** 0.2
        h = h * min(5.0, max(0.2, fac))
    return y, nfe


def divergence_exact(f, y, t):
    """(f(y, t), sum_i d f_i / d y_i) with one autograd backward per dimension; y (R, D)."""
    with torch.enable_grad():
        yy = y.detach().requires_grad_(True)
        v = f(yy, t)
        div = torch.zeros(y.shape[0], dtype=y.dtype)
        for i in range(y.shape[1]):
            g, = torch.autograd.grad(v[:, i].sum(), yy, retain_graph=True)
            div = div + g[:, i]
    return v.detach(), div.detach()


def cnf_log_prob(velocity, theta, t_min, t_max, mean_base, std_base, atol=1e-6, rtol=1e-5):
    """log-density of theta under the flow of `velocity(y (R, D), t) -> (R, D)`:
    integrate [theta, 0] from t_min to t_max with exact trace, add the base log-density."""
    R, D = theta.shape

    def aug(state, t):
        y = state[:R * D].reshape(R, D)
        v, div = divergence_exact(velocity, y, t)
        return torch.cat([v.reshape(-1), div])

    y0 = torch.cat([theta.reshape(-1), torch.zeros(R, dtype=theta.dtype)])
    y, nfe = odeint_dopri5(aug, y0, t_min, t_max, atol=atol, rtol=rtol)
    z, ladj = y[:R * D].reshape(R, D), y[R * D:]
    base = (-0.5 * ((z - mean_base) / std_base) ** 2 - torch.log(std_base) - 0.5 * math.log(2 * math.pi)).sum(1)
    return base + ladj, nfe
