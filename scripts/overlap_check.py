"""A/B of the side lane (SSR_OVERLAP) at the benchmarked configuration, in ONE process on one box:
(1) first-step gradients of every tensor, all parts overlapped vs single-stream vs a second single-stream run (the run-to-run floor);
(2) the captured step timed with CUDA events for every variant ("0", "1", "bwd", ... separated by "/"), interleaved, three rounds of
20 replays.   python scripts/overlap_check.py [B] [num_block] [variants] > out.json"""
import json
import sys

import torch

sys.path.insert(0, ".")
from satlas_super_resolution_b200 import weights
from satlas_super_resolution_b200.trainer import ESRGANTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 23


def rel_l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def make(overlap, graph):
    """overlap: False / True / '0' / '1' / 'bwd' / 'fwd', optionally ':<vgg split>' ('1:all', '1:pool2', ...)"""
    split = None
    if isinstance(overlap, str) and ":" in overlap:
        overlap, split = overlap.split(":")
    return ESRGANTrainer(weights.rrdbnet_state(24, 3, num_block=NB, seed=0), weights.unet_disc_state(27, seed=1), weights.vgg19_state(seed=2),
                         dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=NB), cuda_graph=graph, overlap=overlap, vgg_split=split))


g = torch.Generator().manual_seed(0)
lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)

first = []
TIMING_ONLY = len(sys.argv) > 4 and sys.argv[4] == "timing"
for overlap in (() if TIMING_ONLY else (False, True, False)):
    tr = make(overlap, False)
    p0, d0 = tr.gbuf.flat.clone(), tr.dbuf.flat.clone()
    tr.feed_data(lr, hr)
    tr.optimize_parameters(1)
    torch.cuda.synchronize()
    first.append(({k: v.clone() for k, v in tr.g_grads().items()}, {k: v.clone() for k, v in tr.d_grads().items()},
                  dict(tr.get_current_log()), tr.gbuf.flat - p0, tr.dbuf.flat - d0, tr.gema.flat - p0))
    del tr
out = {"B": B, "num_block": NB}
for name, i in (() if TIMING_ONLY else (("adam_g_update", 3), ("adam_d_update", 4), ("ema_update", 5))):
    out[name + "_overlap_vs_single"] = rel_l2(first[1][i], first[0][i])
    out[name + "_single_vs_single"] = rel_l2(first[2][i], first[0][i])
for name, which in (() if TIMING_ONLY else (("g", 0), ("d", 1))):
    dev = {k: rel_l2(first[1][which][k], v) for k, v in first[0][which].items()}
    flo = {k: rel_l2(first[2][which][k], v) for k, v in first[0][which].items()}
    kd, kf = max(dev, key=dev.get), max(flo, key=flo.get)
    out[f"{name}_grad_overlap_vs_single_worst"] = [kd, dev[kd]]
    out[f"{name}_grad_single_vs_single_worst"] = [kf, flo[kf]]
    out[f"{name}_grad_tensors"] = len(dev)
if not TIMING_ONLY:
    out["losses_single"] = first[0][2]
    out["losses_overlap"] = first[1][2]

variants = sys.argv[3].split("/") if len(sys.argv) > 3 else ["0", "1", "1:all", "1:pool2", "1:pool4", "bwd"]
trs = {v: make(v, True) for v in variants}
for tr in trs.values():
    for it in range(1, 6):
        tr.feed_data(lr, hr)
        tr.optimize_parameters(it)
torch.cuda.synchronize()
times = {v: [] for v in variants}
for rep in range(3):
    for v, tr in trs.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(20):
            tr.optimize_parameters(6 + rep * 20 + it)
        e1.record()
        torch.cuda.synchronize()
        times[v].append(e0.elapsed_time(e1) / 20)
out["ms_per_step"] = times
out["modes"] = {v: tr._last_mode for v, tr in trs.items()}
print(json.dumps(out))
