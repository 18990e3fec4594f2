"""1-D neural spline flow (`posterior_nn("nsf")` with scalar theta / `likelihood_nn("nsf")` with scalar x:
flow.py:401-432, ContextSplineMap :1419-1478) against the UNMODIFIED reference builder through oracle.ref_shim:
same seed -> bit-identical state_dict, including the weight-shared hidden layers."""
import warnings

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")


@pytest.mark.parametrize("C,hl", [(3, 1), (5, 2), (2, 0)])
def test_nsf_1d_builder_matches_reference_builder_bitwise(C, hl):
    assert ref_shim.install()
    from sbi.neural_nets import posterior_nn as ref_posterior_nn
    from sbi_b200.neural_nets import posterior_nn
    g = torch.Generator().manual_seed(0)
    theta = 0.7 * torch.randn(300, 1, generator=g) + 0.3
    x = 1.3 * torch.randn(300, C, generator=g) - 0.2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(5)
        ref = ref_posterior_nn("nsf", hidden_layers_spline_context=hl)(theta, x)
        torch.manual_seed(5)
        est = posterior_nn("nsf", hidden_layers_spline_context=hl)(theta, x)
    want, got = ref.state_dict(), est.state_dict()
    assert set(want) == set(got), set(want) ^ set(got)
    for k in want:
        assert torch.equal(want[k].float(), got[k].float().cpu()), k
    est.load_state_dict(want)
    assert est.layout.NB == hl and est.layout.T == 5
