"""Pins of the oracle's (oracle/nflows_port) remaining arithmetic against independent library code in this image,
where no nflows golden vectors exist: the MADE mixture-of-Gaussians density against
`torch.distributions.MixtureSameFamily`, the autoregressive structure of MADE (both block types) against its
autograd Jacobian, the LU linear transform against `torch.linalg`, the standard-normal base against
`torch.distributions.Normal`, the MAF affine transform against a per-dimension change of variables.  (Spline:
tests/test_oracle_spline_pin_cpu.py; ODE: tests/test_ode_port_cpu.py.)"""
import math

import pytest
import torch
from torch.distributions import Categorical, MixtureSameFamily, Normal
from torch.nn import functional as F

from oracle.nflows_port.distributions import StandardNormal
from oracle.nflows_port.nn.nde import MixtureOfGaussiansMADE
from oracle.nflows_port.transforms import LULinear, MaskedAffineAutoregressiveTransform
from oracle.nflows_port.transforms.made import MADE


def test_made_mog_log_prob_equals_torch_mixture():
    torch.manual_seed(0)
    net = MixtureOfGaussiansMADE(features=4, hidden_features=24, context_features=3, num_blocks=2,
                                 num_mixture_components=5, use_residual_blocks=True, activation=F.relu,
                                 custom_initialization=True).double()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.3 * torch.randn_like(p) * (p != 0))       # keep masked entries masked
    x, ctx = torch.randn(64, 4, dtype=torch.float64), torch.randn(64, 3, dtype=torch.float64)
    out = net(x, context=ctx).reshape(64, 4, 5, 3)
    logits, means, raw = out[..., 0], out[..., 1], out[..., 2]
    mix = MixtureSameFamily(Categorical(logits=logits), Normal(means, F.softplus(raw) + net.epsilon))
    want = mix.log_prob(x).sum(-1)
    assert torch.allclose(net.log_prob(x, context=ctx), want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("residual,act", [(True, F.relu), (False, torch.tanh)])
def test_made_is_autoregressive(residual, act):
    """output block i depends on inputs j < i only (strictly lower-triangular Jacobian).  The context enters through
    the hidden units, and no hidden unit feeds the FIRST feature's outputs (degree 1 needs hidden degree < 1): they
    are constants -- the reason sbi's MADEMoGWrapper prepends a dummy feature (nn_utils.py:133-201)."""
    torch.manual_seed(1)
    D, mult = 5, 3
    net = MADE(features=D, hidden_features=32, context_features=2, num_blocks=2, output_multiplier=mult,
               use_residual_blocks=residual, activation=act).double()
    x, ctx = torch.randn(D, dtype=torch.float64), torch.randn(2, dtype=torch.float64)
    J = torch.autograd.functional.jacobian(lambda v: net(v[None], context=ctx[None])[0], x).reshape(D, mult, D)
    for i in range(D):
        assert bool((J[i, :, i:] == 0).all()), i
        if i > 0:
            assert bool((J[i, :, :i].abs().sum() > 0))
    Jc = torch.autograd.functional.jacobian(lambda c: net(x[None], context=c[None])[0], ctx)
    Jc = Jc.reshape(D, mult, 2)
    assert bool((Jc[0] == 0).all()) and bool((Jc[1:].abs().sum(-1) > 0).all())


def test_lu_linear_equals_dense_linear_algebra():
    torch.manual_seed(2)
    D = 6
    lu = LULinear(D, identity_init=True).double()
    with torch.no_grad():
        for p in lu.parameters():
            p.add_(0.4 * torch.randn_like(p))
    x = torch.randn(32, D, dtype=torch.float64)
    y, ld = lu(x)
    W = lu.weight()
    assert torch.allclose(y, x @ W.T + lu.bias, atol=1e-12)
    assert torch.allclose(ld, torch.linalg.slogdet(W)[1].expand(32), atol=1e-12)
    back, ld_inv = lu.inverse(y)
    assert torch.allclose(back, torch.linalg.solve(W, (y - lu.bias).T).T, atol=1e-10)
    assert torch.allclose(back, x, atol=1e-10) and torch.allclose(ld_inv, -ld, atol=1e-12)
    # unit lower-triangular L, diag(U) = softplus(raw) + eps > 0
    L_, U_ = lu._create_lower_upper()
    assert torch.equal(torch.diagonal(L_), torch.ones(D, dtype=torch.float64)) and bool((torch.diagonal(U_) > 0).all())
    assert torch.equal(L_, torch.tril(L_)) and torch.equal(U_, torch.triu(U_))


def test_standard_normal_equals_torch_normal():
    base = StandardNormal((7,))
    z = torch.randn(50, 7, dtype=torch.float64)
    assert torch.allclose(base.log_prob(z), Normal(0.0, 1.0).log_prob(z).sum(-1), atol=1e-12)


def test_maf_affine_transform_is_the_per_dimension_change_of_variables():
    """z_i = scale_i(x_<i) x_i + shift_i(x_<i): log|det| = sum_i log scale_i, with the reference's scale
    parameterisation softplus(s) + 1e-3; inverse recovers x dimension by dimension."""
    torch.manual_seed(3)
    D = 4
    t = MaskedAffineAutoregressiveTransform(features=D, hidden_features=16, context_features=2, num_blocks=2,
                                            use_residual_blocks=False, activation=torch.tanh).double()
    with torch.no_grad():
        for p in t.parameters():
            p.add_(0.3 * torch.randn_like(p) * (p != 0))
    x, ctx = torch.randn(20, D, dtype=torch.float64), torch.randn(20, 2, dtype=torch.float64)
    z, ld = t(x, context=ctx)
    for r in range(3):
        J = torch.autograd.functional.jacobian(lambda v: t(v[None], context=ctx[r:r + 1])[0][0], x[r])
        assert torch.equal(J, torch.tril(J))
        assert torch.allclose(torch.log(torch.diagonal(J)).sum(), ld[r], atol=1e-12)
    back, ld_inv = t.inverse(z, context=ctx)
    assert torch.allclose(back, x, atol=1e-10) and torch.allclose(ld_inv, -ld, atol=1e-10)
    assert math.isfinite(float(ld.detach().sum()))
