"""Coverage diagnostics on a trained NPE posterior: batched sampling / log_prob of the
DirectPosterior, sbi_b200.diagnostics.run_sbc / run_tarp, and the UNMODIFIED reference's run_sbc + check_sbc
driving the same posterior (drop-in)."""
import math
import warnings

import pytest
import torch
from torch.distributions import MultivariateNormal

from oracle import ref_shim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained(cuda_lib):
    from sbi_b200.inference import NPE
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), torch.eye(D))
    theta = prior.sample((20_000,))
    x = theta + math.sqrt(0.3) * torch.randn_like(theta)
    inf = NPE(prior, density_estimator="nsf", device="cuda")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inf.append_simulations(theta, x).train(training_batch_size=1000, max_num_epochs=60)
    return inf.build_posterior(), prior


def test_sample_batched_and_log_prob_batched(trained):
    post, prior = trained
    torch.manual_seed(1)
    th = prior.sample((64,))
    xs = th + math.sqrt(0.3) * torch.randn_like(th)
    s = post.sample_batched((2000,), x=xs)
    assert s.shape == (2000, 64, 3) and torch.isfinite(s).all()
    want_mean = xs / 1.3
    err = (s.mean(0).cpu() - want_mean).abs()              # the learnt posterior is least accurate for tail observations
    assert err.mean() < 0.05 and err.max() < 0.4
    rel = (s.std(0).cpu() / math.sqrt(0.3 / 1.3) - 1).abs()
    assert rel.mean() < 0.1 and rel.max() < 0.35
    one = post.sample((2000,), x=xs[5:6]).cpu()
    assert (one.mean(0) - s[:, 5].mean(0).cpu()).abs().max() < 0.08
    lp = post.log_prob_batched(s[:10], xs, norm_posterior=False)
    assert lp.shape == (10, 64)
    direct = post.log_prob(s[:10, 7], x=xs[7:8], norm_posterior=False)
    assert (lp[:, 7] - direct).abs().max() < 1e-4
    lpn = post.log_prob_batched(s[:10], xs, norm_posterior=True)
    assert (lpn >= lp - 1e-6).all()           # dividing by an acceptance rate <= 1


def test_map_recovers_posterior_mode(trained):
    post, _ = trained
    x_o = torch.tensor([[0.6, -0.4, 0.2]])
    m = post.map(x=x_o, num_iter=200, num_init_samples=500, num_to_optimize=50).cpu()
    assert (m.reshape(-1) - x_o[0] / 1.3).abs().max() < 0.25        # the mode of a learnt spline density is noisy (posterior std 0.48)


def test_sbc_and_tarp_on_a_calibrated_posterior(trained):
    from scipy.stats import kstest
    from sbi_b200.diagnostics import run_sbc, run_tarp
    post, prior = trained
    torch.manual_seed(2)
    N, S = 400, 500
    th = prior.sample((N,))
    xs = th + math.sqrt(0.3) * torch.randn_like(th)
    ranks, dap = run_sbc(th, xs, post, num_posterior_samples=S)
    assert ranks.shape == (N, 3) and dap.shape == (N, 3)
    for d in range(3):                                            # check_uniformity_frequentist (sbc.py:341-361)
        p = kstest(ranks[:, d].cpu().numpy(), "uniform", args=(0, S))[1]
        assert p > 1e-3, (d, p)
    ecp, alpha = run_tarp(th, xs, post, num_posterior_samples=S)
    mid = alpha.shape[0] // 2
    atc = ((ecp[mid:] - alpha[mid:]) * (alpha[1] - alpha[0])).sum().item()
    assert abs(atc) < 0.05


@pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")
def test_reference_run_sbc_drives_the_b200_posterior(trained):
    assert ref_shim.install()
    from sbi.diagnostics.sbc import check_sbc, run_sbc
    post, prior = trained
    torch.manual_seed(3)
    N, S = 300, 300
    th = prior.sample((N,)).cuda()
    xs = th + math.sqrt(0.3) * torch.randn_like(th)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ranks, dap = run_sbc(th, xs, post, num_posterior_samples=S, show_progress_bar=False)
        stats = check_sbc(ranks.cpu(), th.cpu(), dap.cpu(), num_posterior_samples=S)
    assert (stats["ks_pvals"] > 1e-3).all()
    assert (stats["c2st_ranks"] - 0.5).abs().max() < 0.12
