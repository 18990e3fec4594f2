"""SIMT vs tensor-core training pair over batch sizes (CUDA events, eager launches, L2 flushed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DIM, NUM_SIMS, make_data
from sbi_b200 import _lib as L
from sbi_b200.neural_nets import posterior_nn
lib = L.load()
theta, x = make_data(NUM_SIMS, DIM)
torch.manual_seed(0)
est = posterior_nn("nsf")(theta[:90000], x[:90000]).cuda()
th, xx = theta.cuda(), x.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
m = est._model(nbuf=3)
for B in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    idx = torch.randperm(90000, device="cuda")[:B]
    rows = L.Rows(th.data_ptr(), xx.data_ptr(), idx.data_ptr(), B, 0)
    lp = torch.empty(B, device="cuda"); acc = torch.zeros(2, device="cuda")
    out = []
    for tc in (False, True):
        os.environ["SBI_B200_VJP_TC"] = "1" if tc else "0"
        est._cache.pop("tc_train", None)
        gpart = est._gpart(est.vjp_parts(B))
        run = lambda: est.vjp(m, rows, B, None, -1.0 / B, lp, gpart, None, None, acc)
        for _ in range(3): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); out.append(ts[len(ts) // 2] * 1e3)
    print(f"B={B:6d}  SIMT {out[0]:8.1f} us   tensor-core pair {out[1]:8.1f} us", flush=True)
