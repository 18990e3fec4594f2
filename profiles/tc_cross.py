"""Crossover of the tensor-core vs SIMT NSF kernels at small row counts (sets TC_MIN_ROWS)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import oracle_nsf, b200_from_oracle
flow, theta, x = oracle_nsf(10, 10)
est = b200_from_oracle(flow, theta, x)
g = torch.Generator(device='cuda').manual_seed(0)
xo = x[:1].cuda()
for R in (2048, 4096, 8192, 10000, 16384, 32768):
    th = torch.randn(R, 10, device='cuda', generator=g) * 0.9 + 0.3
    cond = torch.randn(R, 10, device='cuda', generator=g)
    for name, fn in (("logprob shared", lambda: est._logprob_raw(th, xo, True)),
                     ("logprob rowcond", lambda: est._logprob_raw(th, cond, False)),
                     ("sample", lambda: est.inverse_flow(th, xo))):
        res = []
        for tc in ("0", "1"):
            os.environ["SBI_B200_TC"] = tc
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / n * 1e3)
        print(f"R={R:6d} {name:16s} simt {res[0]:8.1f} us   tc {res[1]:8.1f} us", flush=True)
