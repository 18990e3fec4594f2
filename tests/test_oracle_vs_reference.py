"""Cross-check of oracle.sbi_port against the UNMODIFIED reference sbi (imported through
oracle.ref_shim).  Skipped where /root/reference is absent (the GPU box)."""
import warnings

import pytest
import torch

from oracle import ref_shim, sbi_port

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    assert ref_shim.install()
    import sbi  # noqa: F401
    return sbi


@pytest.mark.parametrize("D,C", [(10, 10), (3, 2), (2, 5)])
def test_build_nsf_matches_reference(ref, D, C):
    from sbi.neural_nets import likelihood_nn, posterior_nn
    theta, x = torch.randn(300, D), torch.randn(300, C)
    torch.manual_seed(5)
    a = posterior_nn("nsf")(theta, x)
    torch.manual_seed(5)
    b = sbi_port.build_nsf(theta, x)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    with torch.no_grad():
        assert torch.equal(a.log_prob(theta[:50], x[:50]), b.log_prob(theta[:50], x[:50]))
        assert torch.equal(a.loss(theta[:50], x[:50]), b.loss(theta[:50], x[:50]))
    # NLE swaps the roles (factory.py:316-318)
    torch.manual_seed(6)
    c = likelihood_nn("nsf")(theta, x) if C > 1 else None
    if c is not None:
        torch.manual_seed(6)
        d = sbi_port.build_nsf(x, theta)
        for k in c.state_dict():
            assert torch.equal(c.state_dict()[k], d.state_dict()[k]), k


def test_reference_npe_runs_end_to_end_on_the_port(ref):
    """The reference's own NPE.train / build_posterior / sample / log_prob execute unmodified."""
    from sbi.inference import NPE
    from sbi.neural_nets import posterior_nn
    from torch.distributions import MultivariateNormal
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(3), 0.1 * torch.eye(3))
    theta = prior.sample((600,))
    x = theta + (0.1 ** 0.5) * torch.randn_like(theta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf = NPE(prior, density_estimator=posterior_nn("nsf"), show_progress_bars=False)
        inf.append_simulations(theta, x).train(training_batch_size=100, max_num_epochs=2)
        post = inf.build_posterior()
        s = post.sample((50,), x=x[:1], show_progress_bars=False)
        lp = post.log_prob(s, x=x[:1])
    assert s.shape == (50, 3) and torch.isfinite(lp).all()
