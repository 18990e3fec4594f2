"""Device check of the factorised iid score and one fnpe sampling run (not a pytest test): the same estimator on the
GPU (kernels) and on the CPU in float64 (oracle network), same theta / x_o / t."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.distributions import MultivariateNormal
from tests.test_score_gpu import _pair
from sbi_b200.flowmatching import factorised_iid_score, sample_sde
from sbi_b200.posteriors import prior_to_device
for sde in ("ve", "vp"):
    est, chk, port, theta, x = _pair(sde)
    prior = MultivariateNormal(torch.zeros(4), 2.0 * torch.eye(4))
    th, xo = theta[:200], x[:5]
    t = torch.tensor(0.37)
    with torch.no_grad():
        want = factorised_iid_score(chk, MultivariateNormal(torch.zeros(4, dtype=torch.float64), 2.0 * torch.eye(4, dtype=torch.float64)),
                                    th.double(), xo.double(), t.double())
        got = factorised_iid_score(est, prior_to_device(prior, "cuda"), th.cuda(), xo.cuda(), t.cuda()).cpu().double()
    err = (got - want).abs().max().item() / max(1.0, want.abs().max().item())
    s = sample_sde(est, 500, xo.cuda(), steps=50, iid_method="fnpe", prior=prior_to_device(prior, "cuda"),
                   corrector="langevin", corrector_params=dict(step_size=1e-3, num_steps=2))
    print(f"fnpe {sde}: score rel err {err:.2e}; samples {tuple(s.shape)} finite={bool(torch.isfinite(s).all())} "
          f"mean {s.mean(0).tolist()}")
    assert err < 3e-3
print("fnpe device check ok")
