"""Multi-round training on the GPU: the NPE-C atomic loss through the CUDA log-prob / VJP
kernels against the oracle port on the CPU (same weights, same contrastive draw), two rounds of NPE on the
linear-Gaussian task against the analytic posterior, and the NRE-A / BNRE / NRE-C trainers."""
import math
import warnings

import pytest
import torch
from torch.distributions import MultivariateNormal

from sbi_b200 import multiround as mr
from tests.helpers import b200_from_oracle, oracle_nsf

pytestmark = pytest.mark.gpu


def test_atomic_loss_and_gradient_match_oracle(cuda_lib):
    D, C, B, A = 5, 7, 300, 10
    flow, theta, x = oracle_nsf(D, C, n=1000)
    est = b200_from_oracle(flow, theta, x)
    prior = MultivariateNormal(torch.zeros(D), 2.0 * torch.eye(D))
    th, xx = theta[:B].float(), x[:B].float()
    masks = (torch.arange(B) % 2 == 0).float()
    choices = mr.contrastive_choices(B, A - 1, "cpu")
    flow64 = flow.double()
    flow64.zero_grad()
    prior64 = MultivariateNormal(torch.zeros(D, dtype=torch.float64), 2.0 * torch.eye(D, dtype=torch.float64))

    class _Wrap:                                    # the port's flow behind the estimator interface
        condition_shape = torch.Size([C])

        @staticmethod
        def log_prob(inp, cond):
            return flow64.log_prob(inp, cond)

    want = mr.atomic_log_prob_proposal_posterior(_Wrap, prior64, th.double(), xx.double(), masks.double(), A, True,
                                                 choices=choices)
    (-want.mean()).backward()
    ref_grad = est.layout.pack({k: p.grad for k, p in flow64.named_parameters()}).double()

    from sbi_b200.posteriors import prior_to_device
    est.zero_grad()
    got = mr.atomic_log_prob_proposal_posterior(est, prior_to_device(prior, "cuda"), th.cuda(), xx.cuda(),
                                                masks.cuda(), A, True, choices=choices.cuda())
    (-got.mean()).backward()
    assert (got.detach().cpu().double() - want.detach()).abs().max() < 2e-3
    g = est.flat.grad.cpu().double()
    scale = ref_grad.abs().max()
    assert (g - ref_grad).abs().max() <= 4e-3 * scale, ((g - ref_grad).abs().max() / scale).item()


def test_two_round_npe_linear_gaussian(cuda_lib):
    """Round 1 from the prior, round 2 from the round-1 posterior at x_o (atomic loss): the final
    posterior matches the analytic one (the reference's multi-round acceptance,
    tests/linearGaussian_snpe_test.py:312-372)."""
    from sbi_b200.inference import NPE
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), torch.eye(D))
    x_o = torch.tensor([[0.6, -0.4, 0.2]])
    sim = lambda th: th + math.sqrt(0.3) * torch.randn_like(th)
    inf = NPE(prior, density_estimator="nsf", device="cuda")
    proposal = prior
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in range(2):
            theta = proposal.sample((3000,)).cpu().reshape(-1, D)
            x = sim(theta)
            inf.append_simulations(theta, x, proposal=None if r == 0 else proposal)
            inf.train(num_atoms=10, training_batch_size=200, max_num_epochs=40 if r == 0 else 25)
            posterior = inf.build_posterior()
            proposal = posterior.set_default_x(x_o)
    assert inf._round == 1 and len(inf.summary["epochs_trained"]) == 2
    s = posterior.sample((5000,), x=x_o).cpu()
    # analytic: N(x_o / 1.3, (0.3 / 1.3) I)
    assert (s.mean(0) - x_o[0] / 1.3).abs().max() < 0.08
    assert (s.std(0) / math.sqrt(0.3 / 1.3) - 1).abs().max() < 0.2


@pytest.mark.parametrize("algo", ["NRE_A", "BNRE", "NRE_C"])
def test_nre_variants_train_and_rank_the_posterior(cuda_lib, algo):
    """The trained ratio puts more mass where the analytic posterior does: E_post[log r] > E_prior[log r],
    and the loss decreased."""
    from sbi_b200 import inference
    D = 3
    torch.manual_seed(0)
    prior = MultivariateNormal(torch.zeros(D), torch.eye(D))
    theta = prior.sample((4000,))
    x = theta + math.sqrt(0.3) * torch.randn_like(theta)
    inf = getattr(inference, algo)(prior, classifier="resnet", device="cuda")
    kw = dict(training_batch_size=200, max_num_epochs=15)
    if algo == "BNRE":
        kw["regularization_strength"] = 10.0
    if algo == "NRE_C":
        kw.update(num_classes=5, gamma=1.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est = inf.append_simulations(theta, x).train(**kw)
    vl = inf.summary["validation_loss"]
    assert vl[-1] < vl[0]
    x_o = torch.tensor([[0.6, -0.4, 0.2]])
    post = MultivariateNormal(x_o[0] / 1.3, (0.3 / 1.3) * torch.eye(D)).sample((2000,)).cuda()
    pri = prior.sample((2000,)).cuda()
    xo = x_o.cuda().expand(2000, -1).contiguous()
    with torch.no_grad():
        assert est(post, xo).mean() > est(pri, xo).mean() + 0.5
