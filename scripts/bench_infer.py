"""Inference throughput probe (BASELINE config 5): 16x16-chunk tiles through the 8-frame generator, MPix/s of SR output."""
import json
import sys
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import weights
from satlas_super_resolution_b200.archs import SSR_RRDBNet
from satlas_super_resolution_b200.infer import infer_grid

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = SSR_RRDBNet(24, 3)
net.load_state_dict(weights.rrdbnet_state(24, 3, seed=0))
net = net.cuda().eval()
lr = torch.randint(1, 256, (256, 24, 32, 32), dtype=torch.uint8).pin_memory()
lr_dev = lr.cuda()
for src, tag in ((lr_dev, "resident"), (lr, "e2e_host_pinned")):
    for _ in range(3):
        infer_grid(net, src, batch=batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        c = infer_grid(net, src, batch=batch)
        if tag != "resident":
            c = c.cpu()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"metric": "infer MPix/s (16x16-chunk 2048^2 tile, 8-frame RRDBNet-23)", "mode": tag, "batch": batch,
                      "ms_per_tile": ms, "mpix_per_s": 4.194304 / (ms / 1e3), "tflops": 9.405e12 / (ms / 1e3) / 1e12}))
