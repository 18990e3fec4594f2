"""`build_made` against the UNMODIFIED reference builder (flow.py:37-112 -> MADEMoGWrapper on the nflows port,
through oracle.ref_shim): same seed -> bit-identical state_dict (weights incl. the custom initialisation, masks,
degrees, z-score buffers); the reference state_dict loads verbatim and exports identically."""
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")


@pytest.mark.parametrize("D,C", [(3, 2), (1, 4), (6, 5)])
def test_made_builder_matches_reference_builder_bitwise(D, C):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    from sbi_b200.neural_nets import posterior_nn
    g = torch.Generator().manual_seed(0)
    theta = 0.7 * torch.randn(300, D, generator=g) + 0.3
    x = 1.3 * torch.randn(300, C, generator=g) - 0.2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(5)
        ref = ref_posterior_nn("made")(theta, x)
        torch.manual_seed(5)
        est = posterior_nn("made")(theta, x)
    want, got = ref.state_dict(), est.state_dict()
    assert set(want) == set(got), set(want) ^ set(got)
    for k in want:
        assert torch.equal(want[k].float(), got[k].float().cpu()), k
    # perturb the reference, load, export: exact round trip (masked-out raw weights are kept aside)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    est.load_state_dict(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(v.float(), est.state_dict()[k].float().cpu()), k
    assert est.layout.D == D + 1 and est.layout.M == 10 and est.layout.NB == 5
