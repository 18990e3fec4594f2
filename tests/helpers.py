"""Shared helpers for the parity tests (oracle side = oracle/, CUDA side = sbi_b200)."""
import torch

from oracle import sbi_port


def oracle_nsf(D=10, C=10, n=2000, seed=0, perturb=0.1, lu_perturb=0.1, **kw):
    """Oracle NSF (reference builder restated) with weights moved off their init so that
    every code path (GLU, LU off-diagonals, all spline bins) is exercised."""
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    flow = sbi_port.build_nsf(theta, x, **kw)
    with torch.no_grad():
        for name, p in flow.named_parameters():
            s = lu_perturb if ("entries" in name or "diag" in name) else perturb
            p.add_(s * torch.randn(p.shape, generator=g))
    return flow, theta, x


def b200_from_oracle(flow, theta, x, device="cuda", **kw):
    from sbi_b200.neural_nets import build_nsf
    est = build_nsf(theta, x, **kw)
    est.load_state_dict(flow.state_dict())
    return est.to(device)


def oracle_maf(D=3, C=2, n=2000, seed=0, perturb=0.1, scale_fn="softplus", **kw):
    from oracle.nflows_port.transforms import autoregressive as _ar
    _ar.MAF_SCALE_FN = scale_fn
    g = torch.Generator().manual_seed(seed)
    theta = 0.7 * torch.randn(n, D, generator=g) + 0.3
    x = 1.3 * torch.randn(n, C, generator=g) - 0.2
    torch.manual_seed(seed)
    flow = sbi_port.build_maf(theta, x, **kw)
    with torch.no_grad():
        for name, p in flow.named_parameters():
            p.add_(perturb * torch.randn(p.shape, generator=g))
    return flow, theta, x


def b200_maf_from_oracle(flow, theta, x, device="cuda", scale_fn="softplus", **kw):
    from sbi_b200.neural_nets import build_maf
    est = build_maf(theta, x, maf_scale_softplus=(scale_fn == "softplus"), **kw)
    est.load_state_dict(flow.state_dict())
    return est.to(device)
