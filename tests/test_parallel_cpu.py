"""World-size-2 gloo tests (CPU) of the data-parallel protocol: gradient all-reduce ==
single-process gradient of the union batch, sharded fixed-budget rejection == single-process accept
set and order, shard ranges, parameter broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sbi_port
from sbi_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    out = {}
    # ---- gradient protocol: per-rank batch, upstream -1/(B*world), all-reduce(sum)
    theta, x = sbi_port.linear_gaussian_data(256, 4, seed=0)
    torch.manual_seed(5)
    flow = sbi_port.build_nsf(theta, x, hidden_features=16, num_transforms=2)
    B = 64
    lo, hi = rank * B, (rank + 1) * B
    flow.zero_grad()
    lp = flow.log_prob(theta[lo:hi], x[lo:hi])[0]
    (lp * (-1.0 / (B * world))).sum().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in flow.parameters()])
    parallel.allreduce_flat_gradient(flat)
    out["grad"] = flat
    # ---- parameter broadcast
    w = torch.full((5,), float(rank))
    parallel.broadcast_parameters(w, 0)
    out["bcast"] = w
    # ---- sharded fixed-budget rejection
    target = torch.distributions.MultivariateNormal(torch.zeros(2), 0.2 * torch.eye(2))
    prop = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    acc, idx = parallel.rejection_fixed_budget(
        target.log_prob, lambda n, g: torch.randn(n, 2, generator=g), prop.log_prob,
        log_bound=float(target.log_prob(torch.zeros(2)) - prop.log_prob(torch.zeros(2))), num_proposals=5000, seed=11)
    out["acc"], out["idx"] = acc, idx
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _single():
    theta, x = sbi_port.linear_gaussian_data(256, 4, seed=0)
    torch.manual_seed(5)
    flow = sbi_port.build_nsf(theta, x, hidden_features=16, num_transforms=2)
    flow.zero_grad()
    (-flow.log_prob(theta[:128], x[:128])[0].mean()).backward()
    grad = torch.cat([p.grad.reshape(-1) for p in flow.parameters()])
    target = torch.distributions.MultivariateNormal(torch.zeros(2), 0.2 * torch.eye(2))
    prop = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
    acc, idx = parallel.rejection_fixed_budget(
        target.log_prob, lambda n, g: torch.randn(n, 2, generator=g), prop.log_prob,
        log_bound=float(target.log_prob(torch.zeros(2)) - prop.log_prob(torch.zeros(2))), num_proposals=5000, seed=11)
    return grad, acc, idx


def test_world2_gloo_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    grad1, acc1, idx1 = _single()
    for r in (0, 1):
        assert torch.allclose(res[r]["grad"], grad1, atol=1e-6, rtol=1e-5)
        assert torch.equal(res[r]["bcast"], torch.zeros(5))
        assert torch.equal(res[r]["idx"], idx1) and torch.equal(res[r]["acc"], acc1)
    assert 0 < idx1.numel() < 5000 and (idx1[1:] > idx1[:-1]).all()


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gradient_exchange_is_off_without_a_process_group():
    """Single process: no peer exchange object, callers keep the plain path (no GPU touched)."""
    assert parallel.world() == (0, 1)
    assert parallel.make_gradient_exchange(1000) is None
    assert parallel.shard_range(10, 0, 1) == (0, 10)
