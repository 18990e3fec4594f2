"""NRE classifier logits at one x_o: tensor-core vs SIMT kernel (cfg5 shape: D_theta = D_x = 10)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sbi_b200.ratio import build_resnet_classifier
g = torch.Generator().manual_seed(0)
theta, x = torch.randn(2000, 10, generator=g), torch.randn(2000, 10, generator=g)
est = build_resnet_classifier(theta, x).cuda()
gd = torch.Generator(device='cuda').manual_seed(0)
for R in (10000, 1 << 17, 1 << 20):
    th = torch.randn(R, 10, device='cuda', generator=gd)
    xo = x[:1].cuda()
    for tc in ("0", "1"):
        os.environ["SBI_B200_TC"] = tc
        for _ in range(3): est.logits_raw(th, xo, x_shared=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n): est.logits_raw(th, xo, x_shared=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"ratio R={R} tc={tc}: {ms:.3f} ms  {R/ms/1e3:.1f} M pairs/s", flush=True)
