#!/usr/bin/env python
"""Headline benchmark: NSF-NPE training samples/s (+ posterior log_prob evals/s) on the
linear-Gaussian workload of BASELINE.json configs[1]:
    posterior_nn("nsf"), dim 10, 100 000 sims, training batch 4096, fp32, 1..8 x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one optimisation step (fused forward+backward kernel -> partial-gradient reduce ->
[all-reduce] -> clip+Adam kernel) on one batch of 4096 rows gathered from the HBM-resident
simulation set.  Scaling is weak: every GPU trains on its own 4096-row batch per step and one
gradient all-reduce joins them.  `value` = rows of all ranks / device time (CUDA events, max
over ranks); `e2e` = the same step through the host-buffer C-ABI call
(`sbi_b200_nsf_train_step_host`) with pinned host batches, H2D/D2H inside the timed region.
`--impl reference` times the reference's CPU path (oracle port: DataLoader + nflows-port flow +
clip + Adam, all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 10
NUM_SIMS = 100_000
BATCH = 4096
LOGPROB_ROWS = 1 << 22        # theta rows per log_prob step per GPU (168 MB > L2)
METRIC = "NSF-NPE train samples/sec + posterior log_prob evals/sec @1/2/4/8 GPU"


def make_data(num_sims, dim, seed=0):
    """mini-sbibm gaussian_linear: theta ~ N(0, 0.1 I), x = theta + sqrt(0.1) eps."""
    import torch
    g = torch.Generator().manual_seed(seed)
    theta = math.sqrt(0.1) * torch.randn(num_sims, dim, generator=g)
    x = theta + math.sqrt(0.1) * torch.randn(num_sims, dim, generator=g)
    return theta, x


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons via NVML while the timed regions run."""

    def __init__(self, index=0, period=0.02):
        self.period, self.index = period, index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1)
        loaded = sorted(m for m, u in self.samples if u > 0) or sorted(m for m, _ in self.samples)
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ reference arm
def cpu_reference_train(steps, warmup, max_seconds=None, threads=None):
    """The reference's CPU training step (oracle port of sbi's loop on the nflows port):
    DataLoader(SubsetRandomSampler, drop_last) batch of 4096 -> loss -> backward ->
    clip_grad_norm_(5) -> Adam.  The intra-op thread count is the best of a quick probe over
    {16, 32, 64, all cores} (the tiny ATen ops of this path get slower with too many threads).
    Returns (samples/s, seconds per step, threads used, steps done)."""
    import torch
    from torch.nn.utils.clip_grad import clip_grad_norm_
    from oracle import sbi_port
    cores = os.cpu_count() or 1
    theta, x = make_data(NUM_SIMS, DIM)
    torch.manual_seed(0)
    tr = sbi_port.ReferenceTrainer(sbi_port.build_nsf)
    train_loader, _ = tr.get_dataloaders(theta, x, training_batch_size=BATCH)
    net = sbi_port.build_nsf(theta[tr.train_indices], x[tr.train_indices])
    tr.net = net
    opt = torch.optim.Adam(list(net.parameters()), lr=5e-4)
    it = [iter(train_loader)]

    def one_step():
        t0 = time.perf_counter()
        try:
            batch = next(it[0])
        except StopIteration:
            it[0] = iter(train_loader)
            batch = next(it[0])
        opt.zero_grad()
        losses = tr._losses(batch)
        loss = torch.mean(losses)
        losses.sum().item()
        loss.backward()
        clip_grad_norm_(net.parameters(), max_norm=5.0)
        opt.step()
        return time.perf_counter() - t0

    if threads is None:
        cands = sorted({min(c, cores) for c in (16, 32, 64, cores)})
        best = None
        for c in cands:
            torch.set_num_threads(c)
            one_step()
            dt = one_step()
            if best is None or dt < best[0]:
                best = (dt, c)
        threads = best[1]
    torch.set_num_threads(threads)
    done, t_timed = 0, 0.0
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        dt = one_step()
        if i >= warmup:
            done += 1
            t_timed += dt
        if max_seconds is not None and time.perf_counter() - t_begin > max_seconds and done >= 2:
            break
    sps = done * BATCH / t_timed
    return sps, t_timed / done, threads, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, sec, cores, done = cpu_reference_train(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "linearGaussian NPE nsf dim=10 100k sims batch=4096 (configs[1])",
                   "global_batch": BATCH, "device": "cpu"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{done} optimisation steps of 4096 rows incl. DataLoader collation"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ b200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from sbi_b200 import _lib as L
    from sbi_b200 import build as _build
    from sbi_b200.neural_nets import posterior_nn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    _build.build()
    lib = L.load()

    K, W = args.steps, max(args.warmup, 3)
    theta, x = make_data(NUM_SIMS, DIM, seed=rank)          # every rank its own shard (weak scaling)
    torch.manual_seed(0)
    n_train = int(0.9 * NUM_SIMS)
    est = posterior_nn("nsf")(theta[:n_train], x[:n_train]).to(dev)
    if world > 1:
        dist.broadcast(est.flat.data, 0)
    lay = est.layout
    P = lay.n_params
    theta_d, x_d = theta.to(dev), x.to(dev)
    B = BATCH
    n_part = lib.sbi_b200_nsf_vjp_parts(B)
    gpart = est._gpart(n_part)
    grad = torch.zeros(P, device=dev)
    state = torch.zeros(2 * P, device=dev)
    step_ctr = torch.zeros(2, dtype=torch.int32, device=dev)
    loss_acc = torch.zeros(2, device=dev)
    sumsq = torch.zeros(lib.sbi_b200_sumsq_blocks(P), device=dev)
    mask = est.net._mask
    idx_pool = torch.stack([torch.randperm(n_train, device=dev)[:B] for _ in range(16)])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # N > 1: the flat gradients are summed over NVLink peer memory by our own kernel (csrc/peer.cu);
    # SBI_B200_NCCL=1 keeps the NCCL all-reduce instead (no CUDA graph then)
    peer = None
    grad_local = grad
    if world > 1:
        from sbi_b200.parallel import make_gradient_exchange
        peer = make_gradient_exchange(P)          # None -> NCCL all-reduce
        if peer is not None:
            grad_local = torch.zeros(P, device=dev)
    launches = {"n": 0}

    def train_step(i):
        m = est._model(nbuf=3)
        idx = idx_pool[i % idx_pool.shape[0]]
        rows = L.Rows(theta_d.data_ptr(), x_d.data_ptr(), idx.data_ptr(), B, 0)
        L.check(lib.sbi_b200_nsf_vjp(C.byref(m), C.byref(rows), None, -1.0 / B, None, L.ptr(gpart),
                                     None, None, L.ptr(loss_acc), L.stream_ptr()), "vjp")
        if peer is not None:
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad_local), L.stream_ptr()), "reduce")
            peer.sum(grad_local, grad, mask, sumsq, step_ctr)
            L.check(lib.sbi_b200_adam_clip_step_norm(L.ptr(est.flat.data), L.ptr(grad), L.ptr(state),
                                                     L.ptr(step_ctr), L.ptr(mask), P, 5e-4, 0.9, 0.999, 1e-8,
                                                     5.0, 1.0 / world, L.ptr(sumsq), peer.n_sumsq, L.stream_ptr()), "adam")
        elif world > 1:
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad), L.stream_ptr()), "reduce")
            dist.all_reduce(grad)
            L.check(lib.sbi_b200_adam_clip_step(L.ptr(est.flat.data), L.ptr(grad), L.ptr(state),
                                                L.ptr(step_ctr), L.ptr(mask), P, 5e-4, 0.9, 0.999, 1e-8,
                                                5.0, 1.0 / world, L.stream_ptr()), "adam")
        else:
            L.check(lib.sbi_b200_reduce_partials_norm(L.ptr(gpart), n_part, P, L.ptr(grad), L.ptr(mask),
                                                      L.ptr(sumsq), L.stream_ptr()), "reduce")
            L.check(lib.sbi_b200_adam_clip_step_norm(L.ptr(est.flat.data), L.ptr(grad), L.ptr(state),
                                                     L.ptr(step_ctr), L.ptr(mask), P, 5e-4, 0.9, 0.999, 1e-8,
                                                     5.0, 1.0, L.ptr(sumsq), sumsq.shape[0], L.stream_ptr()), "adam")
        launches["n"] += 4 if peer is not None else 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(index=local)
    clocks.start()

    # ---- training: device-resident leg (value) ---------------------------------------------
    graphs = None
    for i in range(W):
        train_step(i)
    torch.cuda.synchronize()
    if world == 1 or peer is not None:
        if world > 1:
            dist.barrier()
        # one CUDA graph per index slot so that replays carry no host launch gaps
        graphs = []
        for i in range(idx_pool.shape[0]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                train_step(i)
            graphs.append(g)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches["n"] = 0
    barrier()
    for i in range(K):
        flush.zero_()                      # untimed: evict L2 between timed steps
        ev[i][0].record()
        if graphs is not None:
            graphs[i % len(graphs)].replay()
            launches["n"] += 4 if peer is not None else 3
        else:
            train_step(i)
        ev[i][1].record()
    barrier()
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_step = sum(ms_steps) / K
    t = torch.tensor([ms_step], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    train_sps = world * B / (ms_step * 1e-3)
    n_launch = launches["n"]

    # ---- dominant kernel alone (roofline): fused fwd+bwd kernel ------------------------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    m = est._model(nbuf=3)
    for i in range(K):
        flush.zero_()
        idx = idx_pool[i % idx_pool.shape[0]]
        rows = L.Rows(theta_d.data_ptr(), x_d.data_ptr(), idx.data_ptr(), B, 0)
        kev[i][0].record()
        L.check(lib.sbi_b200_nsf_vjp(C.byref(m), C.byref(rows), None, -1.0 / B, None, L.ptr(gpart),
                                     None, None, None, L.stream_ptr()), "vjp")
        kev[i][1].record()
    torch.cuda.synchronize()
    vjp_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    real_params = lay.num_real_params()
    alg_bytes = B * (4 * DIM + 4 * DIM + 4) + 2 * 4 * real_params   # rows in + weights in + grads out
    flops = 2.0 * B * 3 * _nsf_macs(lay)                              # fwd + 2x bwd (no recompute counted)
    peaks = _peaks()
    achieved = alg_bytes / (vjp_ms * 1e-3) / 1e9
    traffic = _traffic()

    # ---- log_prob leg (secondary metric) ---------------------------------------------------------
    R = LOGPROB_ROWS
    th_eval = (math.sqrt(0.1) * torch.randn(R, DIM, device=dev))
    x_o = x_d[:1].contiguous()
    for _ in range(3):
        est._logprob_raw(th_eval, x_o, True)
    lk = max(5, min(K, 20))
    lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(lk)]
    barrier()
    for i in range(lk):
        lev[i][0].record()
        est._logprob_raw(th_eval, x_o, True)
        lev[i][1].record()
    barrier()
    lp_ms = sum(a.elapsed_time(b) for a, b in lev) / lk
    t = torch.tensor([lp_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lp_ms = float(t.item())
    lp_eps = world * R / (lp_ms * 1e-3)
    lp_bytes = R * (4 * DIM + 4)
    lp_flops = 2.0 * R * _nsf_macs(lay)
    tc_used = est._tc_state(est._model(nbuf=2)) is not None and R >= est.TC_MIN_ROWS

    # ---- end to end: host buffers through the C ABI ----------------------------------------------
    ws = L.TrainWs()
    st_in = torch.empty(max(B, 1 << 20), DIM, device=dev)
    st_c = torch.empty(max(B, 1 << 20), DIM, device=dev)
    st_lp = torch.empty(max(B, 1 << 20), device=dev)
    ws.d_input, ws.d_cond, ws.d_logp = st_in.data_ptr(), st_c.data_ptr(), st_lp.data_ptr()
    ws.d_gpart, ws.d_grad, ws.d_state = gpart.data_ptr(), grad.data_ptr(), state.data_ptr()
    ws.d_step, ws.d_mask, ws.d_loss_acc = step_ctr.data_ptr(), mask.data_ptr(), loss_acc.data_ptr()
    ws.cap_rows = st_in.shape[0]
    ws.d_sumsq = sumsq.data_ptr()
    h_th = torch.empty(B, DIM).pin_memory()
    h_x = torch.empty(B, DIM).pin_memory()
    h_th2 = [torch.empty(B, DIM).pin_memory() for _ in range(2)]
    h_x2 = [torch.empty(B, DIM).pin_memory() for _ in range(2)]
    h_loss = torch.zeros(2).pin_memory()
    pipe = lib.sbi_b200_pipe_create()
    perm_host = torch.randperm(n_train)
    e2e = None

    def host_step(i):
        """One optimisation step from a pinned host batch.  N=1: the single blocking C-ABI call
        `sbi_b200_nsf_train_step_host` (H2D, kernels, D2H inside).  N>1: the same pieces with the
        gradient all-reduce between reduce and Adam (H2D / D2H still inside the step)."""
        idx = perm_host[(i * B) % (n_train - B):][:B]
        mm = est._model(nbuf=3)
        if world == 1:
            # pipelined C-ABI step: enqueue step i (H2D + kernels + D2H), get step i-1's loss back
            a, b = h_th2[i & 1], h_x2[i & 1]
            torch.index_select(theta, 0, idx, out=a)      # host batch assembly (the reference's
            torch.index_select(x, 0, idx, out=b)          # DataLoader collation)
            L.check(lib.sbi_b200_nsf_train_step_host_async(
                C.byref(mm), C.byref(ws), pipe, a.data_ptr(), b.data_ptr(), B, 5e-4, 0.9, 0.999, 1e-8,
                5.0, h_loss.data_ptr(), L.stream_ptr()), "train_step_host_async")
            return
        torch.index_select(theta, 0, idx, out=h_th)
        torch.index_select(x, 0, idx, out=h_x)
        st_in[:B].copy_(h_th, non_blocking=True)
        st_c[:B].copy_(h_x, non_blocking=True)
        loss_acc.zero_()
        rows = L.Rows(st_in.data_ptr(), st_c.data_ptr(), None, B, 0)
        L.check(lib.sbi_b200_nsf_vjp(C.byref(mm), C.byref(rows), None, -1.0 / B, None, L.ptr(gpart), None, None,
                                     L.ptr(loss_acc), L.stream_ptr()), "vjp")
        if peer is not None:
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad_local), L.stream_ptr()), "reduce")
            peer.sum(grad_local, grad, mask, sumsq, step_ctr)
            L.check(lib.sbi_b200_adam_clip_step_norm(L.ptr(est.flat.data), L.ptr(grad), L.ptr(state), L.ptr(step_ctr),
                                                     L.ptr(mask), P, 5e-4, 0.9, 0.999, 1e-8, 5.0, 1.0 / world,
                                                     L.ptr(sumsq), peer.n_sumsq, L.stream_ptr()), "adam")
        else:
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad), L.stream_ptr()), "reduce")
            dist.all_reduce(grad)
            L.check(lib.sbi_b200_adam_clip_step(L.ptr(est.flat.data), L.ptr(grad), L.ptr(state), L.ptr(step_ctr),
                                                L.ptr(mask), P, 5e-4, 0.9, 0.999, 1e-8, 5.0, 1.0 / world,
                                                L.stream_ptr()), "adam")
        h_loss.copy_(loss_acc, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    torch.set_num_threads(min(8, os.cpu_count() or 1))   # host-side gathers are tiny: avoid a 128-thread fork/join
    for i in range(W):
        host_step(i)
    if world == 1:
        lib.sbi_b200_pipe_drain(pipe, h_loss.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_step(W + i)
    if world == 1:
        lib.sbi_b200_pipe_drain(pipe, h_loss.data_ptr())   # the last step's loss is read inside the timed region
    barrier()
    e2e_s = (time.perf_counter() - t0) / K
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # log_prob e2e on 2^20 host rows per GPU
    Rh = 1 << 20
    h_eval = th_eval[:Rh].cpu().pin_memory()
    h_xo = x_o.cpu().pin_memory()
    h_out = torch.empty(Rh).pin_memory()
    mm = est._model(nbuf=2)
    tcs = est._tc_state(mm)

    def lp_host():
        if tcs is not None:
            L.check(lib.sbi_b200_nsf_logprob_host_tc(C.byref(mm), C.byref(tcs), C.byref(ws), h_eval.data_ptr(),
                                                     h_xo.data_ptr(), Rh, 1, h_out.data_ptr(), L.stream_ptr()),
                    "logprob_host_tc")
        else:
            L.check(lib.sbi_b200_nsf_logprob_host(C.byref(mm), C.byref(ws), h_eval.data_ptr(), h_xo.data_ptr(),
                                                  Rh, 1, h_out.data_ptr(), L.stream_ptr()), "logprob_host")

    for _ in range(2):
        lp_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        lp_host()
    barrier()
    lp_e2e_s = (time.perf_counter() - t0) / 5
    t = torch.tensor([lp_e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lp_e2e_s = float(t.item())
    e2e = {"value": world * B / e2e_s, "unit": "samples/s", "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": world * B * 2 * DIM * 4, "d2h_bytes_per_step": world * 8,
           "api": "sbi_b200_nsf_train_step_host_async (C ABI, pinned host batch; each step's H2D/D2H inside, "
                  "result of step i read while step i+1 runs)" if world == 1 else
                  ("host batch -> H2D -> vjp -> reduce -> peer-memory gradient sum (csrc/peer.cu) -> clip+Adam -> D2H "
                   "loss, per rank" if peer is not None else
                   "host batch -> H2D -> vjp -> reduce -> NCCL all-reduce -> clip+Adam -> D2H loss, per rank"),
           "log_prob": {"value": world * Rh / lp_e2e_s, "unit": "evals/s", "rows": world * Rh,
                        "api": "sbi_b200_nsf_logprob_host_tc" if tcs is not None else "sbi_b200_nsf_logprob_host",
                        "h2d_bytes_per_step": world * (Rh * DIM * 4 + DIM * 4), "d2h_bytes_per_step": world * Rh * 4}}
    clk = clocks.stop()

    # ---- CPU baseline (bounded sample), rank 0 at N=1 only ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, sec, cores, done = cpu_reference_train(steps=40, warmup=2, max_seconds=20)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{done} optimisation steps of 4096 rows (oracle port of the reference loop incl. "
                         f"DataLoader collation), {sec * 1e3:.0f} ms/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": train_sps, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "linearGaussian NPE posterior_nn('nsf') dim=10 100k sims batch=4096 "
                                   "(BASELINE configs[1]); step = fwd+bwd+clip+Adam on one batch",
                       "global_batch": B * world, "per_gpu_batch": B, "params": real_params,
                       "parallelism": f"dp{world}", "l2": "flushed between timed steps (256 MiB memset, untimed)",
                       "launch": "cuda-graph per step" if graphs is not None else "eager",
                       "gradient_exchange": ("none" if world == 1 else
                                             "peer-memory sum kernel over NVLink (csrc/peer.cu)" if peer is not None
                                             else "NCCL all-reduce")},
            "roofline": {"bound": "hbm", "kernel": "nsf_vjp_kernel<32,2,2,true>", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": traffic.get("nsf_vjp_kernel", {}).get("dram_bytes_per_launch"),
                         "traffic_source": "profiles/r01_traffic.json (ncu --set full, B=4096)",
                         "peak_source": peaks["source"],
                         "alg_bytes_per_launch": alg_bytes, "kernel_ms": vjp_ms,
                         "fp32_fma": {"achieved_tflops": flops / (vjp_ms * 1e-3) / 1e12,
                                      "nominal_peak_tflops": 74.5,
                                      "note": "fused kernel is FP32-FMA/latency bound (SURVEY 8d)"}},
            "secondary": {"metric": "posterior log_prob evals/sec", "value": lp_eps, "unit": "evals/s",
                          "rows_per_step_per_gpu": R, "ms_per_step": lp_ms,
                          "l2": "inputs (168 MB) larger than L2",
                          "roofline": _lp_roofline(tc_used, lp_bytes, lp_flops, lp_ms, peaks)},
            "cpu_baseline": cpu, "clocks": clk, "e2e": e2e, "gpu_launches": n_launch,
            "step_ms_minmax": [min(ms_steps), max(ms_steps)],
        }
        print(json.dumps(line))
    if peer is not None:
        if peer.error():
            raise RuntimeError("peer gradient exchange timed out")
        peer.close()
    if world > 1:
        dist.destroy_process_group()


def _lp_roofline(tc_used, lp_bytes, lp_flops, lp_ms, peaks):
    """Roofline entry of the log_prob kernel.  Tensor-core path: algorithmic fp32-equivalent flops
    (2 x MACs of the linears, no padding, counted once although 3xTF32 issues three MMAs per
    product) against the tf32 tensor peak, taken as half the measured dense bf16 peak."""
    sec = lp_ms * 1e-3
    if not tc_used:
        return {"bound": "hbm", "kernel": "nsf_logprob_kernel<64,4>", "achieved": lp_bytes / sec / 1e9,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": lp_bytes / sec / 1e9 / peaks["hbm_gbs"],
                "fp32_fma_tflops": lp_flops / sec / 1e12}
    peak = peaks["bf16_tflops"] / 2.0
    ach = lp_flops / sec / 1e12
    return {"bound": "tensor", "kernel": "nsf_logprob_tc_kernel<50,10,false>", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": peaks["source"] + "; tf32 = bf16/2",
            "note": "3xTF32: the tensor pipe executes 3x the algorithmic flops (plus K/N padding 50->56/64)",
            "hbm_gbs": lp_bytes / sec / 1e9}


def _nsf_macs(lay):
    """Multiply-accumulates per row of one forward pass (conditioners + LU; spline excluded)."""
    tot = 0
    for l in range(lay.T):
        n_id, n_tr = len(lay.id_feats[l]), len(lay.tr_feats[l])
        tot += (n_id + lay.C) * lay.H + lay.NB * (2 * lay.H * lay.H + lay.C * lay.H)
        tot += lay.H * n_tr * lay.NPAR + lay.D * lay.D
    return tot


def _traffic():
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
