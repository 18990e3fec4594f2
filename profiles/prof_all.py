"""Launch every kernel family of the library once inside a cudaProfiler range (for ncu captures).

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof_all python profiles/prof_all.py

Workload shapes are the BASELINE configs: cfg1 MAF (D=3), cfg2 NSF (D=10, B=4096), cfg4 FMPE (D=20,
B=16384), cfg5 NRE-B (D=10; 2000 training pairs, 2^20 rejection pairs), cfg3 slice chains (1000 x 2-D).
Everything is warmed up outside the range first.
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sbi_b200 import _lib as L  # noqa: E402
from sbi_b200.flowmatching import _FmLoss, posterior_flow_nn  # noqa: E402
from sbi_b200.neural_nets import posterior_nn  # noqa: E402
from sbi_b200.ratio import classifier_nn  # noqa: E402
from sbi_b200.samplers import SliceSamplerVectorized  # noqa: E402

only = set(sys.argv[1:])
lib = L.load()
torch.manual_seed(0)
jobs = []


def job(name):
    def deco(fn):
        if not only or name in only:
            jobs.append((name, fn))
        return fn
    return deco


def gauss(n, d):
    th = math.sqrt(0.1) * torch.randn(n, d)
    return th, th + math.sqrt(0.1) * torch.randn(n, d)


# ---- cfg2: NSF ------------------------------------------------------------------------------------
th10, x10 = gauss(20000, 10)
nsf = posterior_nn("nsf")(th10, x10).cuda()
th10c, x10c = th10.cuda(), x10.cuda()


@job("nsf_train")
def nsf_train():
    nsf.zero_grad()
    nsf.loss(th10c[:4096], x10c[:4096]).mean().backward()


@job("nsf_eval")
def nsf_eval():
    with torch.no_grad():
        nsf.log_prob(th10c[:512], x10c[:512])                       # SIMT kernel
        nsf.log_prob(th10c.unsqueeze(1), x10c[:1])                  # tensor-core kernel (20000 rows)
        nsf.sample((512,), x10c[:1])
        nsf.sample((20000,), x10c[:1])


# ---- cfg1: MAF ------------------------------------------------------------------------------------
th3, x3 = gauss(20000, 3)
maf = posterior_nn("maf")(th3, x3).cuda()
th3c, x3c = th3.cuda(), x3.cuda()


@job("maf")
def maf_all():
    maf.zero_grad()
    maf.loss(th3c[:4096], x3c[:4096]).mean().backward()
    with torch.no_grad():
        maf.log_prob(th3c.unsqueeze(1), x3c[:1])
        maf.sample((20000,), x3c[:1])


# ---- cfg5: NRE-B resnet classifier ----------------------------------------------------------------
rat = classifier_nn("resnet")(th10, x10).cuda()
th1m = math.sqrt(0.1) * torch.randn(1 << 20, 10, device="cuda")


@job("ratio")
def ratio_all():
    rat.zero_grad()
    t = th10c[:2000].clone().requires_grad_(True)
    rat(t, x10c[:2000]).sum().backward()
    with torch.no_grad():
        rat(th10c[:4096], x10c[:4096])                              # SIMT forward
        rat.logits_raw(th1m, x10c[:1].contiguous(), x_shared=True)  # tensor-core forward, 2^20 pairs at x_o


# ---- cfg4: FMPE -----------------------------------------------------------------------------------
th20, x20 = gauss(32768, 20)
fm = posterior_flow_nn("mlp")(th20, x20).cuda()
th20c, x20c = th20.cuda(), x20.cuda()
tt = torch.rand(16384, device="cuda")
ee = torch.randn(16384, 20, device="cuda")


@job("fm")
def fm_all():
    fm.zero_grad()
    _FmLoss.apply(fm.net.flat, th20c[:16384], x20c[:16384], tt, ee, fm).mean().backward()
    with torch.no_grad():
        fm.forward(th20c[:16384], x20c[:1], torch.tensor(0.3, device="cuda"))


# ---- cfg3: slice sampler state machine ------------------------------------------------------------
def _logp(p):
    return -0.5 * (p * p).sum(-1)


@job("slice")
def slice_all():
    SliceSamplerVectorized(_logp, np.zeros((1000, 2)), num_chains=1000, thin=1, tuning=5, seed=1).run(3)


# ---- optimizer tail -------------------------------------------------------------------------------
P = nsf.layout.n_params
n_part = lib.sbi_b200_nsf_vjp_parts(4096)
gpart = torch.randn(n_part, P, device="cuda") * 1e-3
grad = torch.zeros(P, device="cuda")
state = torch.zeros(2 * P, device="cuda")
step = torch.zeros(2, dtype=torch.int32, device="cuda")


@job("optim")
def optim_all():
    L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad), L.stream_ptr()), "red")
    L.check(lib.sbi_b200_adam_clip_step(L.ptr(nsf.flat.data), L.ptr(grad), L.ptr(state), L.ptr(step),
                                        L.ptr(nsf.net._mask), P, 5e-4, .9, .999, 1e-8, 5.0, 1.0,
                                        L.stream_ptr()), "adam")


for _ in range(2):
    for _, fn in jobs:
        fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _, fn in jobs:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done:", [n for n, _ in jobs])
