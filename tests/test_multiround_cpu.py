"""Loss heads of sbi_b200/multiround.py against the UNMODIFIED reference's own methods (CPU, through
oracle.ref_shim): NPE-C atomic proposal posterior (npe_c.py:356-440), NRE-A, BNRE and NRE-C losses.
The heads are device-agnostic torch code; here they run on the reference's CPU estimator so that the
comparison is exact.  (GPU side: tests/test_multiround_gpu.py.)"""
import warnings

import pytest
import torch
from torch.distributions import MultivariateNormal

from oracle import ref_shim
from sbi_b200 import multiround as mr

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")


@pytest.fixture(scope="module")
def ref():
    assert ref_shim.install()
    import sbi  # noqa: F401
    return sbi


def _data(n=300, D=3, C=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    theta = 0.8 * torch.randn(n, D, generator=g)
    x = torch.cat([theta, theta], 1)[:, :C] + 0.5 * torch.randn(n, C, generator=g)
    return theta, x


@pytest.mark.parametrize("combined", [False, True])
def test_atomic_proposal_posterior_equals_reference(ref, combined):
    from sbi.inference import NPE_C
    from sbi.neural_nets import posterior_nn
    theta, x = _data()
    prior = MultivariateNormal(torch.zeros(3), torch.eye(3))
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf = NPE_C(prior, density_estimator=posterior_nn("nsf"), show_progress_bars=False)
        net = posterior_nn("nsf")(theta, x)
    inf._neural_net, inf._num_atoms, inf._use_combined_loss = net, 7, combined
    B = 64
    th, xx = theta[:B], x[:B]
    masks = (torch.arange(B) % 3 == 0).float().reshape(-1, 1)
    choices = mr.contrastive_choices(B, 6, "cpu")
    assert (choices != torch.arange(B)[:, None]).all() and all(len(set(r.tolist())) == 6 for r in choices)
    real = torch.multinomial
    torch.multinomial = lambda *a, **k: choices
    try:
        want = inf._log_prob_proposal_posterior_atomic(th, xx, masks)
    finally:
        torch.multinomial = real
    got = mr.atomic_log_prob_proposal_posterior(net, prior, th, xx, masks, 7, combined, choices=choices)
    assert torch.equal(got, want)


def test_nre_heads_equal_reference(ref):
    from sbi.inference import BNRE, NRE_A, NRE_C
    B = 50
    g = torch.Generator().manual_seed(1)
    theta, x = _data(B)

    def with_logits(inf, seq):
        it = iter(seq)
        inf._classifier_logits = lambda th, xx, n: next(it).reshape(-1, 1)
        inf._device = "cpu"
        return inf

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        l2 = torch.randn(B, 2, generator=g)
        a = with_logits(NRE_A(show_progress_bars=False), [l2])
        assert torch.allclose(mr.nre_a_loss(l2), a._loss(theta, x, 2), atol=0, rtol=0)
        b = with_logits(BNRE(show_progress_bars=False), [l2])
        assert torch.allclose(mr.bnre_loss(l2, 100.0), b._loss(theta, x, 2, 100.0), atol=0, rtol=0)
        K = 5
        lm, lj = torch.randn(B, K + 1, generator=g), torch.randn(B, K, generator=g)
        c = with_logits(NRE_C(show_progress_bars=False), [lm, lj])
        assert torch.allclose(mr.nre_c_loss(lm, lj, 0.7), c._loss(theta, x, K + 1, 0.7), atol=0, rtol=0)
