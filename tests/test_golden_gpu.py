"""The CUDA kernels against the committed golden fixtures (outputs of the UNMODIFIED reference sbi,
generated in the build container by tests/golden/make_golden.py): reference state_dict loaded
verbatim into the sbi_b200 estimators, reference inputs in, reference outputs expected.
Tolerances: fp32 kernels vs fp32 reference: log-prob / logits 2e-3 abs, samples and velocity
fields 2e-3, per-row FM loss 2e-3 relative."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["nsf_d10.pt", "nsf_d3c2.pt"])
@pytest.mark.parametrize("tc", ["0", "1"])
def test_nsf_kernels_reproduce_reference_fixture(cuda_lib, monkeypatch, name, tc):
    """tc=0: SIMT kernels; tc=1: tensor-core kernels forced (64-row inputs are below their default threshold)."""
    from sbi_b200.neural_nets import build_nsf
    monkeypatch.setenv("SBI_B200_TC", tc)
    g = torch.load(os.path.join(GOLD, name))
    est = build_nsf(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    est = est.cuda()
    inp, cond, noise = g["inp"].cuda(), g["cond"].cuda(), g["noise"].cuda()
    with torch.no_grad():
        lp = est.log_prob(inp, cond)[0].cpu()
        lps = est.log_prob(inp.unsqueeze(1), cond[:1])[:, 0].cpu()
        z = est.inverse_transform(inp, cond).cpu()
        s, lad = est.inverse_flow(noise, cond)
    assert (lp - g["log_prob"]).abs().max() <= 2e-3
    assert (lps - g["log_prob_shared"]).abs().max() <= 2e-3
    assert (z - g["inverse_transform"]).abs().max() <= 2e-3
    assert (s.cpu() - g["samples"]).abs().max() <= 2e-3
    assert (lad.cpu() - g["inverse_logabsdet"]).abs().max() <= 5e-3


def test_maf_kernels_reproduce_reference_fixture(cuda_lib):
    from sbi_b200.neural_nets import build_maf
    g = torch.load(os.path.join(GOLD, "maf_d3c2.pt"))
    est = build_maf(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    est = est.cuda()
    with torch.no_grad():
        lp = est.log_prob(g["inp"].cuda(), g["cond"].cuda())[0].cpu()
        s, lad = est.inverse_flow(g["noise"].cuda(), g["cond"].cuda())
    assert (lp - g["log_prob"]).abs().max() <= 2e-3
    assert (s.cpu() - g["samples"]).abs().max() <= 2e-3
    assert (lad.cpu() - g["inverse_logabsdet"]).abs().max() <= 5e-3


@pytest.mark.parametrize("tc", ["0", "1"])
def test_ratio_kernels_reproduce_reference_fixture(cuda_lib, monkeypatch, tc):
    from sbi_b200.ratio import build_resnet_classifier
    monkeypatch.setenv("SBI_B200_TC", tc)
    g = torch.load(os.path.join(GOLD, "ratio_d4x6.pt"))
    est = build_resnet_classifier(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    est = est.cuda()
    with torch.no_grad():
        a = est(g["th"].cuda(), g["xx"].cuda()).cpu()
        b = est.logits_raw(g["th"].cuda(), g["xx"][:1].cuda(), x_shared=True).cpu()
    assert (a - g["logits"]).abs().max() <= 2e-3
    assert (b - g["logits_shared"]).abs().max() <= 2e-3


def test_flow_matching_kernels_reproduce_reference_fixture(cuda_lib):
    from sbi_b200.flowmatching import build_vector_field_estimator
    g = torch.load(os.path.join(GOLD, "fm_d5c3.pt"))
    est = build_vector_field_estimator(g["theta"], g["x"])
    est.load_state_dict(g["state_dict"])
    est = est.cuda()
    inp, cond, t = g["inp"].cuda(), g["cond"].cuda(), g["t"].cuda()
    with torch.no_grad():
        v = est(inp, cond, t).cpu()
        vs = est(inp, cond[:1], torch.tensor(0.37, device="cuda")).cpu()
        loss = est.loss_raw(inp.contiguous(), cond.contiguous(), t.contiguous(), g["theta_1"].cuda().contiguous())
        loss = (loss[0] if isinstance(loss, tuple) else loss).cpu()
    assert (v - g["v"]).abs().max() <= 2e-3
    assert (vs - g["v_shared"]).abs().max() <= 2e-3
    assert ((loss - g["loss"]).abs() / g["loss"].abs().clamp_min(1e-3)).max() <= 2e-3
