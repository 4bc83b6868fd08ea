"""Host-side ingest throughput: the reference's per-sample PNG decode vs the packed uint8 shard.

Synthetic S2-NAIP tree (N chips, T Sentinel-2 frames per chip, TCI only = the 8-frame RGB config), one process, one thread,
samples drawn in order.  Prints one JSON line; `python scripts/bench_ingest.py [N] [T]`.
"""
import json
import os
import random
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from satlas_super_resolution_b200.data import PinnedBatcher, S2NAIPShardDataset, pack_s2naip  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.set_num_threads(1)
root = tempfile.mkdtemp(prefix="ssr_ingest_")
rng = np.random.default_rng(0)
for i in range(N):
    chip = f"{1000 + i}_{2000 + i}"
    os.makedirs(os.path.join(root, "naip", "2020", chip))
    os.makedirs(os.path.join(root, "s2", chip))
    cv2.imwrite(os.path.join(root, "naip", "2020", chip, chip + ".png"), rng.integers(1, 256, (128, 128, 3), dtype=np.uint8))
    cv2.imwrite(os.path.join(root, "s2", chip, "tci.png"), rng.integers(1, 256, (T * 32, 32, 3), dtype=np.uint8))
opt = dict(phase="train", n_s2_images=8, scale=4, sentinel2_path=os.path.join(root, "s2"), naip_path=os.path.join(root, "naip"))
t0 = time.perf_counter()
pack_s2naip(dict(opt), os.path.join(root, "shard"))
t_pack = time.perf_counter() - t0


def rate(ds, batch=32):
    pb = PinnedBatcher(ds, batch, pin=False)
    random.seed(0)
    pb.batch(range(batch))                       # warm the page cache / memmap
    t0 = time.perf_counter()
    n = 0
    for s in range(0, len(ds) - batch + 1, batch):
        pb.batch(range(s, s + batch))
        n += batch
    return n / (time.perf_counter() - t0)


png = rate(S2NAIPShardDataset(dict(opt)))
shard = rate(S2NAIPShardDataset(dict(opt, shard_path=os.path.join(root, "shard"))))
print(json.dumps({"metric": "ingest img-pairs/s per host thread (8 of T frames, RGB, batch 32 into a staging buffer)", "chips": N,
                  "frames_per_chip": T, "png_tree": png, "packed_shard": shard, "speedup": shard / png,
                  "pack_seconds": t_pack, "shard_bytes": os.path.getsize(os.path.join(root, "shard.bin")),
                  "host": f"{os.cpu_count()} logical CPUs (build container, not the B200 host)", "threads": 1}))
