"""sbi_b200.diagnostics rank / coverage statistics against the UNMODIFIED reference's `_run_sbc` / `_run_tarp`
on the same posterior-sample tensors (CPU, exact), and `DirectPosterior.sample_batched`'s acceptance bookkeeping
with a stand-in estimator."""
import pytest
import torch

from oracle import ref_shim
from sbi_b200 import diagnostics as dg

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no copy of the reference sbi")


def test_sbc_ranks_and_tarp_equal_reference():
    assert ref_shim.install()
    from sbi.diagnostics.sbc import _run_sbc
    from sbi.diagnostics.tarp import _run_tarp
    g = torch.Generator().manual_seed(0)
    S, N, D = 200, 150, 3
    thetas = torch.randn(N, D, generator=g)
    xs = thetas + 0.3 * torch.randn(N, D, generator=g)
    samples = xs.unsqueeze(0) + 0.3 * torch.randn(S, N, D, generator=g)
    want = _run_sbc(thetas, xs, samples, "marginals", show_progress_bar=False)
    assert torch.equal(dg.sbc_ranks(thetas, xs, samples, "marginals"), want)
    fn = lambda th, x: (th * x).sum(-1)
    want_fn = _run_sbc(thetas, xs, samples, fn, show_progress_bar=False)
    assert torch.equal(dg.sbc_ranks(thetas, xs, samples, fn), want_fn)
    refs = dg.get_tarp_references(thetas)
    for z in (False, True):
        e0, a0 = _run_tarp(samples, thetas, refs, num_bins=None, z_score_theta=z)
        e1, a1 = dg.tarp_coverage(samples, thetas, refs, num_bins=None, z_score_theta=z)
        assert torch.equal(e0, e1) and torch.equal(a0, a1)
