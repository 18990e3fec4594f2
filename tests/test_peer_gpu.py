"""Gradient sum over NVLink peer memory (csrc/peer.cu, parallel.PeerGradientSum) against NCCL
all-reduce, on two GPUs of one node (skipped on a single-GPU box)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from sbi_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    P = 98025
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    mask = (torch.rand(P, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) > 0.1).to(torch.uint8)
    ex = parallel.PeerGradientSum(P)
    out = torch.zeros(P, device=dev)
    sumsq = torch.zeros(ex.n_sumsq, device=dev)
    step = torch.zeros(2, dtype=torch.int32, device=dev)
    loc = torch.zeros(P, device=dev)
    ok = True
    msgs = []

    def check(tag):
        nonlocal ok
        ref = loc.clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        if not torch.allclose(out, ref, rtol=1e-6, atol=1e-6):
            ok = False
            msgs.append(f"{tag}: sum mismatch {(out - ref).abs().max().item():.3e}")
        want = (out.double() * mask.double()).pow(2).sum().item()
        got = sumsq.double().sum().item()
        if abs(got - want) > 1e-5 * max(want, 1.0):
            ok = False
            msgs.append(f"{tag}: sumsq {got} vs {want}")
        both = [torch.zeros_like(out) for _ in range(world)]
        dist.all_gather(both, out)
        if not all(torch.equal(both[0], b) for b in both):
            ok = False
            msgs.append(f"{tag}: ranks differ bitwise")

    # eager steps
    for s in range(6):
        loc.copy_(torch.randn(P, device=dev, generator=g))
        ex.sum(loc, out, mask, sumsq, step)
        check(f"eager{s}")
        step[0] += 1
    # rewound step counter (warm-up then restore): stale flags must not satisfy the wait
    step[0] = 0
    for s in range(3):
        loc.copy_(torch.randn(P, device=dev, generator=g))
        ex.sum(loc, out, mask, sumsq, step)
        check(f"rewound{s}")
        step[0] += 1
    # CUDA graph replay
    torch.cuda.synchronize()
    dist.barrier()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ex.sum(loc, out, mask, sumsq, step)
        step[0] += 1
    torch.cuda.current_stream().wait_stream(side)
    check("pre-capture")
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ex.sum(loc, out, mask, sumsq, step)
        step[0:1].add_(1)
    for s in range(4):
        loc.copy_(torch.randn(P, device=dev, generator=g))
        graph.replay()
        check(f"graph{s}")
    if ex.error():
        ok = False
        msgs.append("peer wait timed out")
    ex.close()
    q.put((rank, ok, msgs))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_gradient_sum_matches_nccl(cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msgs in res:
        assert ok, (rank, msgs)


def _train_worker(rank, world, port, q):
    import math
    import torch.distributed as dist
    from torch.distributions import MultivariateNormal
    from sbi_b200.inference import NPE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    D = 3
    torch.manual_seed(10 + rank)                       # every rank simulates its own shard
    prior = MultivariateNormal(torch.zeros(D), 0.1 * torch.eye(D))
    theta = prior.sample((4000,))
    x = theta + math.sqrt(0.1) * torch.randn_like(theta)
    torch.manual_seed(0)                               # same split / permutation stream on every rank
    inf = NPE(prior, density_estimator="nsf", device=f"cuda:{rank}").data_parallel()
    est = inf.append_simulations(theta, x).train(training_batch_size=500, max_num_epochs=6)
    flat = est.flat.data.clone()
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    same = all(torch.equal(both[0], b) for b in both)
    vl = inf.summary["validation_loss"]
    q.put((rank, same, [float(v) for v in vl]))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_training_two_gpus(cuda_lib):
    """NPE.data_parallel().train() on two GPUs: peer-memory gradient sum inside the per-epoch CUDA
    graph; the replicas end bit-identical and the validation loss goes down."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, vl in res:
        assert same, f"rank {rank}: replicas diverged"
        assert all(v == v for v in vl) and vl[-1] < vl[0], vl
