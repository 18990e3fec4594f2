"""Pin of the oracle's rational-quadratic spline arithmetic against an INDEPENDENT third-party implementation that
ships in this image: `transformers.models.vits.modeling_vits._unconstrained_rational_quadratic_spline` (the
VITS text-to-speech model carries its own copy of the Durkan et al. 2019 monotone-spline code, the lineage nflows
0.14's `unconstrained_rational_quadratic_spline(tails="linear")` comes from).  nflows itself is absent from the
image and the wheelhouse, so this is the closest external known-answer source for SURVEY row a7:
bin search with the +1e-6 last knot, softmax widths / heights with their floors, softplus derivatives with unit
boundary slopes, the forward map, the quadratic-root inverse, log|det|, identity tails.  The oracle
(`oracle/nflows_port`) is what every CUDA spline kernel is compared with."""
import pytest
import torch

from oracle.nflows_port.transforms.splines import rational_quadratic as port

vits = pytest.importorskip("transformers.models.vits.modeling_vits")


def _case(n, K, seed, dtype, scale=2.0, tail=3.0):
    g = torch.Generator().manual_seed(seed)
    x = ((torch.rand(n, generator=g) * 8.0 - 4.0) * tail / 3.0).to(dtype)    # a quarter of the rows outside the tails
    uw = (scale * torch.randn(n, K, generator=g)).to(dtype)
    uh = (scale * torch.randn(n, K, generator=g)).to(dtype)
    ud = (scale * torch.randn(n, K - 1, generator=g)).to(dtype)
    return x, uw, uh, ud


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)])
@pytest.mark.parametrize("K,tail", [(10, 3.0), (10, 5.0), (4, 3.0), (16, 1.0)])
@pytest.mark.parametrize("inverse", [False, True])
def test_port_equals_the_vits_spline(dtype, tol, K, tail, inverse):
    x, uw, uh, ud = _case(4000, K, seed=K + int(tail), dtype=dtype, tail=tail)
    got, got_ld = port.unconstrained_rational_quadratic_spline(
        x.clone(), uw.clone(), uh.clone(), ud.clone(), inverse=inverse, tails="linear", tail_bound=tail)
    want, want_ld = vits._unconstrained_rational_quadratic_spline(
        x.clone(), uw.clone(), uh.clone(), ud.clone(), reverse=inverse, tail_bound=tail)
    assert torch.allclose(got, want, rtol=tol, atol=tol), (got - want).abs().max()
    assert torch.allclose(got_ld, want_ld, rtol=tol, atol=tol), (got_ld - want_ld).abs().max()
    outside = x.abs() > tail
    assert outside.any() and torch.equal(got[outside], x[outside]) and bool((got_ld[outside] == 0).all())


def test_nsf_parameterisation_matches_through_the_coupling_defaults():
    """The reference's NSF settings (num_bins = 10, tail_bound = 3, min widths / heights / derivative 1e-3): knots
    exactly at a bin boundary and at the tail bound fall into the same bins in both implementations."""
    K, tail = 10, 3.0
    x, uw, uh, ud = _case(512, K, seed=1, dtype=torch.float64)
    # put a block of inputs exactly onto knots of their own spline (forward direction: width knots)
    w = torch.softmax(uw, -1)
    w = 1e-3 + (1 - 1e-3 * K) * w
    knots = 2 * tail * torch.cumsum(w, -1) - tail
    x[:64] = knots[:64, 4]
    x[64:96] = tail
    x[96:128] = -tail
    got = port.unconstrained_rational_quadratic_spline(x.clone(), uw, uh, ud, tails="linear", tail_bound=tail)
    want = vits._unconstrained_rational_quadratic_spline(x.clone(), uw.clone(), uh.clone(), ud.clone(), tail_bound=tail)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-12)
    # round trip through the OTHER implementation: port forward, VITS inverse
    y, ld = got
    back, ld_inv = vits._unconstrained_rational_quadratic_spline(y.clone(), uw.clone(), uh.clone(), ud.clone(),
                                                                 reverse=True, tail_bound=tail)
    assert torch.allclose(back, x, atol=1e-9) and torch.allclose(ld_inv, -ld, atol=1e-9)
