import os, sys, torch, time
sys.path.insert(0, '/root/repo')
from tests.helpers import oracle_nsf, b200_from_oracle
flow, theta, x = oracle_nsf(10, 10)
est = b200_from_oracle(flow, theta, x)
g = torch.Generator(device='cuda').manual_seed(0)
for R in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
    th = torch.randn(R, 10, device='cuda', generator=g) * 0.9 + 0.3
    xo = x[:1].cuda()
    for tc in ("0", "1"):
        os.environ["SBI_B200_TC"] = tc
        for _ in range(3): est._logprob_raw(th, xo, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): est._logprob_raw(th, xo, True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"R={R} tc={tc}: {ms:.3f} ms  {R/ms/1e3:.1f} M evals/s", flush=True)

# sampling direction
for R in (1 << 18, 1 << 20):
    noise = torch.randn(R, 10, device='cuda', generator=g)
    xo = x[:1].cuda()
    for tc in ("0", "1"):
        os.environ["SBI_B200_TC"] = tc
        for _ in range(3): est.inverse_flow(noise, xo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): est.inverse_flow(noise, xo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"sample R={R} tc={tc}: {ms:.3f} ms  {R/ms/1e3:.1f} M samples/s", flush=True)
