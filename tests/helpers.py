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


def two_moons_simulator(parameters, r_loc=0.1, r_scale=0.01, base_offset=0.25):
    """The two-moons simulator of the reference's mini benchmark, restated
    (/root/reference/tests/mini_sbibm/two_moons.py:15-78): a noisy half circle shifted by
    (-|z0|, z1), z = parameters rotated by -45 degrees."""
    import math
    n = parameters.shape[0]
    a = (torch.rand(n, 1) - 0.5) * math.pi
    r = r_loc + r_scale * torch.randn(n, 1)
    p = torch.cat((torch.cos(a) * r + base_offset, torch.sin(a) * r), dim=1)
    c, s = math.cos(-math.pi / 4.0), math.sin(-math.pi / 4.0)
    z0 = (c * parameters[:, 0] - s * parameters[:, 1]).reshape(-1, 1)
    z1 = (s * parameters[:, 0] + c * parameters[:, 1]).reshape(-1, 1)
    return p + torch.cat((-torch.abs(z0), z1), dim=1)


def c2st(X, Y, seed=1, n_folds=5):
    """Classifier two-sample test accuracy as the reference computes it by default
    (/root/reference/sbi/utils/metrics.py:56-190): both sets z-scored with X's statistics, a
    RandomForestClassifier (100 trees... sklearn defaults) scored by 5-fold shuffled cross-validation."""
    import numpy as np
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.model_selection import KFold, cross_val_score
    X, Y = X.double(), Y.double()
    mean, std = X.mean(0), X.std(0)
    std[std == 0] = 1.0
    X, Y = (X - mean) / std, (Y - mean) / std
    data = np.concatenate((X.numpy(), Y.numpy()))
    target = np.concatenate((np.zeros(X.shape[0]), np.ones(Y.shape[0])))
    clf = RandomForestClassifier(random_state=seed, n_jobs=-1)
    shuffle = KFold(n_splits=n_folds, shuffle=True, random_state=seed)
    return float(cross_val_score(clf, data, target, cv=shuffle, scoring="accuracy").mean())
