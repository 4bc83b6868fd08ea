"""Device time of the generator forward / backward plans replayed as CUDA graphs (B pairs, 32x32 tiles)."""
import sys
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import weights
from satlas_super_resolution_b200.generator import RRDBNetEngine
from satlas_super_resolution_b200.ops import cur_stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sd = {k: v.cuda() for k, v in weights.rrdbnet_state(24, 3, seed=0).items()}
grads = {k: torch.zeros_like(v) for k, v in sd.items()}
eng = RRDBNetEngine(sd, 24, 3, want_grad=True, grads=grads)
eng.repack()
x = torch.rand(B, 24, 32, 32, device="cuda")
d_out = torch.randn(B, 3, 128, 128, device="cuda") * 1e-5
eng.forward(x, train=True)
eng.backward(d_out, B, 32, 32)
torch.cuda.synchronize()
ws = eng.workspace(B, 32, 32, True)


def graph_of(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn(cur_stream())
    return g


def timeit(g, n=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


gf = graph_of(lambda s: ws.fwd.run(s))
gb = graph_of(lambda s: ws.bwd.run(s))
tf, tb = timeit(gf), timeit(gb)
FG = 36.739e9 * B
print(f"B={B} G forward  {tf:7.3f} ms  {len(ws.fwd)} launches  {FG / tf / 1e9:7.1f} TFLOP/s")
print(f"B={B} G backward {tb:7.3f} ms  {len(ws.bwd)} launches  {2 * FG / tb / 1e9:7.1f} TFLOP/s (dgrad + wgrad)")
