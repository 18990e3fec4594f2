"""Pins for the oracle itself (CPU): the reference's known-answer test of the path, the golden
fixtures produced by the UNMODIFIED reference, and mathematical self-checks of the nflows port
(the only pins available for the third-party arithmetic: see DESIGN.md §5)."""
import os

import pytest
import torch

from oracle import sbi_port
from oracle.nflows_port.transforms.splines.rational_quadratic import (
    unconstrained_rational_quadratic_spline as urqs,
)
from oracle.nflows_port.utils import torchutils as nf_utils

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("fn", [sbi_port.searchsorted, nf_utils.searchsorted])
def test_searchsorted_known_answer(fn):
    """/root/reference/tests/torchutils_test.py:135-157 (the path's only fixed-vector test)."""
    bin_locations = torch.linspace(0, 1, 10)
    cases = [bin_locations[:-1].clone(), bin_locations[:-1] + 0.1, bin_locations[:-1] + 0.05]
    gold = torch.load(os.path.join(GOLD, "searchsorted.pt"))
    for inputs, name in zip(cases, ("left_boundaries", "right_boundaries", "mid_points")):
        idx = fn(bin_locations[None, :], inputs)
        assert torch.equal(idx, torch.arange(0, 9))
        assert torch.equal(idx, gold[name]["idx"])
    shape = [2, 3, 4]
    idx = fn(torch.linspace(0, 1, 10).repeat(*shape, 1), torch.rand(*shape))
    assert idx.shape == torch.Size(shape)


@pytest.mark.parametrize("name", ["nsf_d10.pt", "nsf_d3c2.pt"])
def test_port_reproduces_reference_fixture(name):
    """oracle.sbi_port (used on the GPU box) == the reference's own outputs, bit for bit: both
    sit on the same nflows port, so any difference is a restatement error in sbi_port."""
    g = torch.load(os.path.join(GOLD, name))
    flow = sbi_port.build_nsf(g["theta"], g["x"])
    flow.load_state_dict(g["state_dict"])
    with torch.no_grad():
        assert torch.equal(flow.log_prob(g["inp"], g["cond"])[0], g["log_prob"])
        assert torch.equal(flow.log_prob(g["inp"].unsqueeze(1), g["cond"][:1])[:, 0], g["log_prob_shared"])
        assert torch.equal(flow.inverse_transform(g["inp"], g["cond"]), g["inverse_transform"])
        emb = flow.net._embedding_net(g["cond"])
        s, lad = flow.net._transform.inverse(g["noise"], context=emb)
        assert torch.equal(s, g["samples"]) and torch.equal(lad, g["inverse_logabsdet"])
    # the builder itself: same seed -> same initial weights and z-score statistics
    torch.manual_seed(g["seed"])
    fresh = sbi_port.build_nsf(g["theta"], g["x"])
    assert set(fresh.state_dict()) == set(g["state_dict"])
    for k in ("net._transform._transforms.0._shift", "net._transform._transforms.0._scale",
              "net._embedding_net.0._mean", "net._embedding_net.0._std"):
        assert torch.equal(fresh.state_dict()[k], g["state_dict"][k])


def test_port_reproduces_reference_maf_fixture():
    g = torch.load(os.path.join(GOLD, "maf_d3c2.pt"))
    flow = sbi_port.build_maf(g["theta"], g["x"])
    flow.load_state_dict(g["state_dict"])
    with torch.no_grad():
        assert torch.equal(flow.log_prob(g["inp"], g["cond"])[0], g["log_prob"])
        s, lad = flow.net._transform.inverse(g["noise"], context=flow.net._embedding_net(g["cond"]))
        assert torch.equal(s, g["samples"]) and torch.equal(lad, g["inverse_logabsdet"])
    torch.manual_seed(g["seed"])
    fresh = sbi_port.build_maf(g["theta"], g["x"])
    for k, v in g["state_dict"].items():
        if k.endswith("_permutation") or k.endswith("mask") or k.endswith("degrees"):
            assert torch.equal(fresh.state_dict()[k], v), k   # same seed -> same permutations / masks


def test_reference_training_trajectory():
    """oracle.sbi_port.ReferenceTrainer == the reference's NPE.train() (same seeds -> same split,
    same validation-loss trajectory, same final weights)."""
    g = torch.load(os.path.join(GOLD, "npe_train.pt"))
    torch.manual_seed(11)
    tr = sbi_port.ReferenceTrainer(sbi_port.build_nsf)
    net = tr.train(g["theta"], g["x"], training_batch_size=200, max_num_epochs=3)
    assert torch.equal(tr.train_indices, g["train_indices"])
    assert tr.summary["validation_loss"] == pytest.approx(g["validation_loss"], rel=1e-6)
    assert tr.summary["training_loss"] == pytest.approx(g["training_loss"], rel=1e-6)
    for k, v in g["state_dict"].items():
        assert torch.allclose(net.state_dict()[k].float(), v.float(), atol=1e-6), k


def _flow(D, C, seed=0):
    g = torch.Generator().manual_seed(seed)
    theta, x = torch.randn(500, D, generator=g), torch.randn(500, C, generator=g)
    torch.manual_seed(seed)
    flow = sbi_port.build_nsf(theta, x)
    with torch.no_grad():
        for p in flow.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    return flow.double(), theta.double(), x.double()


def test_inverse_of_forward_is_identity():
    flow, theta, x = _flow(6, 3)
    emb = flow.net._embedding_net(x[:200])
    z, ld = flow.net._transform(theta[:200] * 1.5, context=emb)
    back, ldi = flow.net._transform.inverse(z, context=emb)
    assert (back - theta[:200] * 1.5).abs().max() < 1e-9
    assert (ld + ldi).abs().max() < 1e-9


def test_logabsdet_matches_autograd_jacobian():
    flow, theta, x = _flow(4, 2)
    for i in range(5):
        ctx = flow.net._embedding_net(x[i:i + 1])
        f = lambda t: flow.net._transform(t[None], context=ctx)[0][0]  # noqa: E731
        J = torch.autograd.functional.jacobian(f, theta[i] * 1.2)
        _, ld = flow.net._transform(theta[i:i + 1] * 1.2, context=ctx)
        assert abs(torch.linalg.slogdet(J)[1].item() - ld.item()) < 1e-8


def test_density_integrates_to_one_2d():
    flow, theta, x = _flow(2, 2)
    n = 801
    # the flow z-scores its input: integrate in input space over +-8 input-std around the mean
    sh = flow.state_dict()["net._transform._transforms.0._shift"]
    sc = flow.state_dict()["net._transform._transforms.0._scale"]
    mean, std = (-sh / sc), 1 / sc
    g0 = torch.linspace(-8, 8, n, dtype=torch.float64)
    gx, gy = torch.meshgrid(mean[0] + std[0] * g0, mean[1] + std[1] * g0, indexing="ij")
    pts = torch.stack([gx.reshape(-1), gy.reshape(-1)], 1)
    with torch.no_grad():
        lp = flow.log_prob(pts.unsqueeze(1), x[:1])[:, 0]
    cell = (16 * std[0] / (n - 1)) * (16 * std[1] / (n - 1))
    assert abs(float(lp.exp().sum() * cell) - 1.0) < 2e-3


def test_spline_is_monotone_and_identity_outside():
    torch.manual_seed(0)
    K = 10
    p = torch.randn(1, 3 * K - 1, dtype=torch.float64) * 3
    xs = torch.linspace(-4, 4, 40001, dtype=torch.float64)
    y, ld = urqs(xs, p[:, :K].expand(40001, K), p[:, K:2 * K].expand(40001, K), p[:, 2 * K:].expand(40001, K - 1),
                 tails="linear", tail_bound=3.0)
    assert (y[1:] > y[:-1]).all()
    out = xs.abs() > 3
    assert torch.equal(y[out], xs[out]) and (ld[out] == 0).all()
    num = (y[2:] - y[:-2]) / (xs[2:] - xs[:-2])
    inside = (xs[1:-1].abs() < 2.9)
    # central difference of y against exp(logabsdet): O(h) only at the C1 knots
    err = (num.log() - ld[1:-1])[inside].abs()
    assert torch.quantile(err, 0.995) < 5e-3   # minimum-width bins (1e-3) are under-resolved by the grid


def test_port_reproduces_reference_ratio_fixture():
    """oracle port of the NRE `resnet` classifier == the reference's own logits (fixture generated by
    tests/golden/make_golden.py from the unmodified reference)."""
    g = torch.load(os.path.join(GOLD, "ratio_d4x6.pt"))
    est = sbi_port.build_resnet_classifier(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    with torch.no_grad():
        assert torch.equal(est(g["th"], g["xx"]), g["logits"])
        assert torch.equal(est(g["th"], g["xx"][:1].expand(96, -1)), g["logits_shared"])


def test_port_reproduces_reference_flow_matching_fixture():
    """oracle port of the flow-matching estimator == the reference's velocity field and loss."""
    g = torch.load(os.path.join(GOLD, "fm_d5c3.pt"))
    est = sbi_port.build_flow_matching_estimator(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    with torch.no_grad():
        assert torch.allclose(est(g["inp"], g["cond"], g["t"]), g["v"], atol=1e-6, rtol=1e-5)
        assert torch.allclose(est(g["inp"], g["cond"][:1], torch.tensor(0.37)), g["v_shared"], atol=1e-6, rtol=1e-5)
        assert torch.allclose(est.loss(g["inp"], g["cond"], times=g["t"], theta_1=g["theta_1"]), g["loss"],
                              atol=1e-6, rtol=1e-5)
