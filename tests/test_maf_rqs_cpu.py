"""`build_maf_rqs` against the UNMODIFIED reference builder (flow.py:212-330 on the nflows port, through
oracle.ref_shim): same seed -> bit-identical state_dict (weights, masks, degrees, permutations, z-score
buffers); reference state_dict loads verbatim."""
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")


@pytest.mark.parametrize("D,C,bins", [(4, 3, 10), (2, 5, 6)])
def test_maf_rqs_builder_matches_reference_builder_bitwise(D, C, bins):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    from sbi_b200.neural_nets import posterior_nn
    g = torch.Generator().manual_seed(0)
    theta = 0.7 * torch.randn(300, D, generator=g) + 0.3
    x = 1.3 * torch.randn(300, C, generator=g) - 0.2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(5)
        ref = ref_posterior_nn("maf_rqs", num_bins=bins)(theta, x)
        torch.manual_seed(5)
        est = posterior_nn("maf_rqs", num_bins=bins)(theta, x)
    want, got = ref.state_dict(), est.state_dict()
    assert set(want) == set(got), set(want) ^ set(got)
    for k in want:
        assert torch.equal(want[k].float(), got[k].float().cpu()), k
    est.load_state_dict(want)
    assert est.layout.OUTM == 3 * bins - 1 and est.layout.head == "rqs"
