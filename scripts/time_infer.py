"""Quick device-time probe of the generator forward (not the bench contract; see bench.py)."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import nets
from satlas_super_resolution_b200.generator import RRDBNetEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = {k: v.cuda() for k, v in nets.rrdbnet_init(24, 3, seed=0).items()}
eng = RRDBNetEngine(p, 24, 3, want_grad=False)
eng.repack()
x = torch.rand(B, 24, 32, 32, device="cuda")
for train in (False, True):
    for _ in range(3):
        eng.forward(x, train=train)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        eng.forward(x, train=train)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 36.739e9 * B
    print(f"train={train} B={B}: {ms:.3f} ms/fwd  {B/ms*1e3:.1f} img/s  {B*0.016384/ms*1e3:.1f} MPix/s  {flops/ms/1e9:.1f} TFLOP/s")
