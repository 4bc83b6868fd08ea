"""Library baseline on the GPU box: the SAME training step / generator forward as the reference would run it on a GPU -- stock
PyTorch ops dispatched to cuDNN / cuBLAS (last row: "the same reference modules on the B200 through stock
PyTorch/cuDNN (.cuda(), TF32 default) as the library baseline the hand-written kernels must beat").

/root/reference does not exist on the GPU box and basicsr is not installable, so the step is the oracle's line-by-line torch
restatement (oracle/step.py, oracle/nets.py -- the same torch calls the reference modules make: F.conv2d, F.interpolate, torch.cat,
F.leaky_relu, BCEWithLogits, torch.optim.Adam) with every tensor moved to cuda:0.  Test / measurement infrastructure only: the
product never imports this.  Three arithmetic settings:

  fp32      torch.backends.cudnn.allow_tf32 = False   (the reference's CPU arithmetic, on the GPU)
  tf32      allow_tf32 = True, cudnn.benchmark = True (what ssr/train.py:34 + torch defaults give the reference on a GPU)
  bf16      tf32 + torch.autocast(bfloat16) + channels_last inputs (the strongest stock-library configuration; NOT the reference's)

usage: python scripts/library_baseline.py [--batch 32] [--steps 5] [--warmup 2] [--out gpurun_out/library_baseline.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def to_dev(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def time_train(mode, B, steps, warmup, dev):
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    gp = to_dev(nets.rrdbnet_init(24, 3, num_block=23, seed=0), dev)
    dp = to_dev(nets.unet_disc_init(27, seed=1), dev)
    vp = to_dev(losses.vgg19_init(seed=2), dev)
    orc = OracleESRGAN(gp, dp, vp, dict(ema_decay=0.999, lr=1e-4), num_block=23)
    g = torch.Generator().manual_seed(3)
    lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8).pin_memory()
    hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8).pin_memory()

    def step(i):
        orc.feed_data(lr.to(dev, non_blocking=True), hr.to(dev, non_blocking=True))
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return orc.optimize_parameters(i)
        return orc.optimize_parameters(i)

    for i in range(warmup):
        step(i + 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        log = step(warmup + i + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(mode=mode, batch=B, ms_per_step=ms, img_pairs_per_s=B / ms * 1e3, steps=steps, warmup=warmup,
                losses={k: float(v) for k, v in log.items()})


def time_infer(mode, chunks, steps, warmup, dev):
    from oracle import nets
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    gp = to_dev(nets.rrdbnet_init(24, 3, num_block=23, seed=0), dev)
    x = torch.rand(chunks, 24, 32, 32, device=dev)
    if mode == "bf16":
        x = x.contiguous(memory_format=torch.channels_last)

    def fwd():
        with torch.no_grad():
            if mode == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return nets.rrdbnet_forward(gp, x)
            return nets.rrdbnet_forward(gp, x)

    for _ in range(warmup):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(mode=mode, chunks=chunks, ms_per_tile=ms, mpix_per_s=chunks * 128 * 128 / 1e6 / ms * 1e3, steps=steps, warmup=warmup)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--chunks", type=int, default=256)
    ap.add_argument("--modes", default="tf32,bf16,fp32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "library_baseline.json"))
    a = ap.parse_args()
    assert torch.cuda.is_available(), "library_baseline.py needs a GPU"
    dev = torch.device("cuda:0")
    res = dict(what="oracle/step.py + oracle/nets.py (the reference's torch calls) on cuda:0 through stock PyTorch / cuDNN / cuBLAS",
               torch=torch.__version__, cudnn=torch.backends.cudnn.version(), gpu=torch.cuda.get_device_name(0),
               when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), train=[], infer=[])
    for mode in a.modes.split(","):
        try:
            r = time_train(mode, a.batch, a.steps, a.warmup, dev)
        except Exception as e:   # report, keep the other modes
            r = dict(mode=mode, error=f"{type(e).__name__}: {e}"[:400])
        print(json.dumps(r), flush=True)
        res["train"].append(r)
        torch.cuda.empty_cache()
        try:
            r = time_infer(mode, a.chunks, a.steps, a.warmup, dev)
        except Exception as e:
            r = dict(mode=mode, error=f"{type(e).__name__}: {e}"[:400])
        print(json.dumps(r), flush=True)
        res["infer"].append(r)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
