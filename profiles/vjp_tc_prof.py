"""The tensor-core training pair (forward sweep with activation save + backward sweep) at the bench
shape, inside a cudaProfiler range for ncu; also prints CUDA-event times of the pair and of the SIMT
kernel.   python profiles/vjp_tc_prof.py [rows]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DIM, NUM_SIMS, make_data
from sbi_b200 import _lib as L
from sbi_b200.neural_nets import posterior_nn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = L.load()
theta, x = make_data(NUM_SIMS, DIM)
torch.manual_seed(0)
est = posterior_nn("nsf")(theta[:90000], x[:90000]).cuda()
th, xx = theta.cuda(), x.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
idx = torch.randperm(90000, device="cuda")[:B]
m = est._model(nbuf=3)
rows = L.Rows(th.data_ptr(), xx.data_ptr(), idx.data_ptr(), B, 0)
lp = torch.empty(B, device="cuda")
acc = torch.zeros(2, device="cuda")


def timed(tc):
    os.environ["SBI_B200_VJP_TC"] = "1" if tc else "0"
    est._cache.pop("tc_train", None)
    gpart = est._gpart(est.vjp_parts(B))
    run = lambda: est.vjp(m, rows, B, None, -1.0 / B, lp, gpart, None, None, acc)
    for _ in range(5): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"B={B} tc={tc}: median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us", flush=True)
    return run


timed(False)
run = timed(True)
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
