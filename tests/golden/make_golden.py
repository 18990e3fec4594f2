"""Generate the golden fixtures from the UNMODIFIED reference sbi (/root/reference), imported on
top of oracle.nflows_port through oracle.ref_shim.  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures (small, committed):
  nsf_d10.pt      reference `posterior_nn("nsf")` build under seed 7 (D=C=10): state_dict (after a
                  deterministic perturbation so all code paths are exercised), inputs, and the
                  reference's log_prob / loss / inverse_transform / sample(noise) outputs.
  nsf_d3c2.pt     same for theta-dim 3, x-dim 2 (the shape of tests/linearGaussian_snpe_test.py:312-372).
  npe_train.pt    a short reference `NPE(...).train()` run (seed 3, 2000 sims, batch 200, 4 epochs):
                  validation-loss trajectory, final state_dict, train/val indices.
  maf_d3c2.pt     reference `posterior_nn("maf")` (theta-dim 3, x-dim 2; BASELINE configs[0]): state_dict,
                  permutations, log_prob and inverse outputs.
  maf_rqs_d4c3.pt reference `posterior_nn("maf_rqs")` (theta-dim 4, x-dim 3): same contents as maf_d3c2.pt.
  nsf_d1c3.pt     the ONE-dimensional NSF (scalar theta, x-dim 3; ContextSplineMap conditioner): as nsf_d3c2.pt.
  made_d3c2.pt    reference `posterior_nn("made")` (theta-dim 3, x-dim 2): state_dict, log_prob values, and
                  moments / quantiles of 20 000 reference samples at one condition.
  searchsorted.pt the reference's bin-search known-answer test vectors (tests/torchutils_test.py:135-157).
  ratio_d4x6.pt   reference `classifier_nn("resnet")` (theta-dim 4, x-dim 6): state_dict, pairs, logits.
  fm_d5c3.pt      reference `posterior_flow_nn("mlp")` (theta-dim 5, x-dim 3): state_dict, inputs, times,
                  velocity field, and the flow-matching loss for given (t, theta_1).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

assert ref_shim.install(), "needs /root/reference"
from sbi.inference import NPE  # noqa: E402
from sbi.neural_nets import posterior_nn  # noqa: E402
from sbi.utils.torchutils import searchsorted  # noqa: E402


def flow_fixture(D, C, seed, n=400):
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    est = posterior_nn("nsf")(theta, x)
    with torch.no_grad():
        for name, p in est.named_parameters():
            s = 0.1
            p.add_(s * torch.randn(p.shape, generator=g))
    inp, cond = theta[:64] * 1.5, x[:64]
    noise = torch.randn(64, D, generator=g)
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0]
        lp_shared = est.log_prob(inp.unsqueeze(1), cond[:1])[:, 0]
        z = est.inverse_transform(inp, cond)
        emb = est.net._embedding_net(cond)
        samples, lad = est.net._transform.inverse(noise, context=emb)
    return dict(state_dict=est.state_dict(), theta=theta, x=x, inp=inp, cond=cond, noise=noise,
                log_prob=lp, log_prob_shared=lp_shared, inverse_transform=z, samples=samples,
                inverse_logabsdet=lad, D=D, C=C, seed=seed)


def maf_fixture(D, C, seed, n=400, model="maf"):
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    est = posterior_nn(model)(theta, x)
    with torch.no_grad():
        for name, p in est.named_parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    inp, cond = theta[:64] * 1.5, x[:64]
    noise = torch.randn(64, D, generator=g)
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0]
        emb = est.net._embedding_net(cond)
        samples, lad = est.net._transform.inverse(noise, context=emb)
    return dict(state_dict=est.state_dict(), theta=theta, x=x, inp=inp, cond=cond, noise=noise,
                log_prob=lp, samples=samples, inverse_logabsdet=lad, D=D, C=C, seed=seed)


def made_fixture(D, C, seed, n=400):
    """reference `posterior_nn("made")`: log_prob at given inputs; `sample` is RNG-bound (Categorical draws),
    so only moments of a large reference sample are stored."""
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    est = posterior_nn("made")(theta, x)
    with torch.no_grad():
        for name, p in est.named_parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    inp, cond = theta[:64] * 1.5, x[:64]
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0]
        lp_shared = est.log_prob(inp.unsqueeze(1), cond[:1])[:, 0]
        torch.manual_seed(seed + 1)
        s = est.sample((20000,), cond[:1])[:, 0]
    return dict(state_dict=est.state_dict(), theta=theta, x=x, inp=inp, cond=cond, log_prob=lp,
                log_prob_shared=lp_shared, sample_mean=s.mean(0), sample_std=s.std(0),
                sample_q=torch.quantile(s, torch.tensor([0.1, 0.5, 0.9]), dim=0), D=D, C=C, seed=seed)


def train_fixture():
    torch.manual_seed(3)
    D = 4
    theta = (0.1 ** 0.5) * torch.randn(2000, D)
    x = theta + (0.1 ** 0.5) * torch.randn(2000, D)
    inf = NPE(density_estimator=posterior_nn("nsf"), show_progress_bars=False)
    torch.manual_seed(11)
    est = inf.append_simulations(theta, x).train(training_batch_size=200, max_num_epochs=3)
    return dict(theta=theta, x=x, validation_loss=inf._summary["validation_loss"],
                training_loss=inf._summary["training_loss"], state_dict=est.state_dict(),
                train_indices=inf.train_indices, val_indices=inf.val_indices)


def ratio_fixture(Dt, Dx, seed, n=400):
    from sbi.neural_nets import classifier_nn
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, Dt, generator=g) + 0.3
    x = 1.3 * torch.randn(n, Dx, generator=g) - 0.2
    torch.manual_seed(seed)
    est = classifier_nn("resnet")(theta, x)
    with torch.no_grad():
        for name, p in est.named_parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
        th, xx = theta[:96] * 1.2, x[:96]
        logits = est(th, xx)
        logits_shared = est(th, xx[:1].expand(96, -1))
    return dict(state_dict=est.state_dict(), theta=theta, x=x, th=th, xx=xx, logits=logits,
                logits_shared=logits_shared, Dt=Dt, Dx=Dx, seed=seed)


def fm_fixture(D, C, seed, n=500):
    from sbi.neural_nets import posterior_flow_nn
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    est = posterior_flow_nn("mlp")(theta, x)
    with torch.no_grad():
        for name, p in est.named_parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
        inp, cond = theta[:64] * 1.1, x[:64]
        t = torch.rand(64, generator=g)
        v = est(inp, cond, t)
        v_shared = est(inp, cond[:1], torch.tensor(0.37))
        torch.manual_seed(seed + 1)
        theta_1 = torch.randn_like(inp)          # the draw est.loss makes right after `times`
        torch.manual_seed(seed + 1)
        loss = est.loss(inp, cond, times=t)
    return dict(state_dict=est.state_dict(), theta=theta, x=x, inp=inp, cond=cond, t=t, v=v,
                v_shared=v_shared, theta_1=theta_1, loss=loss, D=D, C=C, seed=seed)


def searchsorted_fixture():
    """Exactly the cases of /root/reference/tests/torchutils_test.py:135-157, evaluated by the
    reference's own `sbi.utils.torchutils.searchsorted` (expected there: arange(0, 9))."""
    bin_locations = torch.linspace(0, 1, 10)
    cases = {"left_boundaries": bin_locations[:-1].clone(),
             "right_boundaries": bin_locations[:-1] + 0.1,
             "mid_points": bin_locations[:-1] + 0.05}
    out = {"bin_locations": torch.linspace(0, 1, 10)}
    for name, inputs in cases.items():
        out[name] = dict(inputs=inputs.clone(), idx=searchsorted(bin_locations[None, :], inputs))
    return out


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    torch.save(flow_fixture(10, 10, 7), os.path.join(HERE, "nsf_d10.pt"))
    torch.save(flow_fixture(3, 2, 8), os.path.join(HERE, "nsf_d3c2.pt"))
    torch.save(train_fixture(), os.path.join(HERE, "npe_train.pt"))
    torch.save(maf_fixture(3, 2, 9), os.path.join(HERE, "maf_d3c2.pt"))
    if "--new" in sys.argv or not os.path.exists(os.path.join(HERE, "nsf_d1c3.pt")):
        torch.save(flow_fixture(1, 3, 16), os.path.join(HERE, "nsf_d1c3.pt"))
    if "--new" in sys.argv or not os.path.exists(os.path.join(HERE, "made_d3c2.pt")):
        torch.save(made_fixture(3, 2, 15), os.path.join(HERE, "made_d3c2.pt"))
    if "--new" in sys.argv or not os.path.exists(os.path.join(HERE, "maf_rqs_d4c3.pt")):
        torch.save(maf_fixture(4, 3, 14, model="maf_rqs"), os.path.join(HERE, "maf_rqs_d4c3.pt"))
    torch.save(searchsorted_fixture(), os.path.join(HERE, "searchsorted.pt"))
    torch.save(ratio_fixture(4, 6, 12), os.path.join(HERE, "ratio_d4x6.pt"))
    torch.save(fm_fixture(5, 3, 13), os.path.join(HERE, "fm_d5c3.pt"))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
