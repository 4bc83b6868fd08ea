#!/usr/bin/env python
"""bench.py -- img-pairs/s of the 8-frame ESRGAN 4x training step (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" = feed_data kernels (uint8 -> float/255, USM sharpen) + SSRESRGANModel.optimize_parameters
(ssr/models/ssr_esrgan_model.py:104-233: G forward, L1 + VGG19-perceptual + 0.1*GAN losses through the frozen D,
G backward + Adam + EMA, D real/fake forward+backward + Adam) on one synthetic batch of B pairs per GPU
(lr uint8 [B,24,32,32], hr uint8 [B,3,128,128]), random-init weights (no checkpoints offline).
  value : pairs/s, whole job, inputs already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e   : same through the public call sequence with HOST (pinned) uint8 batches: H2D copy inside the timed region,
          loss scalars read back (D2H) every step
  roofline     : the dominant kernel (ssr_conv_tc: every conv forward / input-gradient of G, D, VGG), algorithmic conv
                 FLOPs per step / its summed device time, measured with per-launch CUDA events in one extra eager step
  cpu_baseline : the CPU restatement of the reference step (oracle/step.py, torch fp32) on this box's host cores
--impl reference: the reference's own CPU path for the same step.  The reference (pure Python on basicsr, which is not
installable offline) cannot run here, so this arm times the oracle port of it ("kind": "port") with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# work model, BASELINE.md section 3 (FLOP = 2*MAC, convs only, per img-pair, RGB 8-frame)
F_G, F_D, F_V, F_FIRST, F_C0 = 36.7390e9, 13.4134e9, 12.7402e9, 0.0283e9, 0.5096e9
FLOP_STEP = 3 * F_G - F_FIRST + 8 * F_D - 2 * F_C0 + 3 * F_V           # 254.70 GFLOP
FLOP_WGRAD = F_G + 2 * F_D                                              # weight gradients (ssr_wgrad_tc)
FLOP_CONV_TC = FLOP_STEP - FLOP_WGRAD                                   # forward + input-gradient convs (ssr_conv_tc)


def env_int(name, default):
    return int(os.environ.get(name, default))


def log(msg):
    if os.environ.get("SSR_BENCH_VERBOSE"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def usable_cores():
    """host threads this process may actually use: affinity mask capped by the cgroup CPU quota (os.cpu_count() reports
    the whole host, and oversubscribing a quota-limited container makes OpenMP spin for minutes)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fq, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fp:
                q, per = int(fq.read()), int(fp.read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def _dominant_traffic():
    """DRAM bytes per launch of the kernel that dominates class 0 (the chained launch: 58 % of the conv time, 138 of 239 launches)"""
    t = ncu_traffic()
    e = dict(t.get("conv_chain_kernel") or t.get("conv_tc_kernel") or {})
    if "note" in e and "conv_chain_kernel" in t:
        e["note"] = "conv_chain_kernel: " + e["note"]
    return e


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel, taken from the committed `ncu --set full` capture of the same step
    (profiles/ncu_traffic.json, written by scripts/summarize_ncu.py) -- a profiler-side number, never measured in-run."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops_sustained", 1379.2), d.get("bf16_tflops", 1660.0), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
    hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    return lr, hr


CONFIG = {"workload": "ESRGAN 8-S2-frame RGB training (RRDBNet-23 + UNetDiscriminatorSN, 4x), synthetic 32x32 tiles",
          "losses": "L1(1.0) + VGG19 perceptual(conv1_2..5_4) + 0.1*GAN(vanilla), Adam 1e-4, EMA 0.999, USM gt",
          "num_in_ch_g": 24, "num_in_ch_d": 27}


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_step_time(batch, iters, warm, threads):
    """the oracle restatement of the reference step (torch fp32) on the host cores; returns s/iter"""
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    torch.set_num_threads(threads)
    gp = nets.rrdbnet_init(24, 3, seed=0)
    dp = nets.unet_disc_init(27, seed=1)
    vp = losses.vgg19_init(seed=2)
    orc = OracleESRGAN(gp, dp, vp, dict(ema_decay=0.999, lr=1e-4))
    lr, hr = synthetic_batch(batch, 0)
    times = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        orc.feed_data(lr, hr)
        orc.optimize_parameters()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = usable_cores()
    batch = 2
    steps = max(1, min(args.steps, 3))
    t = cpu_step_time(batch, steps, 1 if args.warmup > 0 else 0, threads)
    val = batch / t
    line = {"impl": "reference", "metric": "img-pairs/sec 8-frame ESRGAN 4x train", "value": val, "unit": "img-pairs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, batch_per_step=batch),
            "cpu_baseline": {"value": val, "unit": "img-pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} full optimize_parameters steps of {batch} pairs (oracle/step.py, torch fp32 CPU); "
                                       "the reference itself needs basicsr, which is absent offline"},
            "e2e": {"value": val, "unit": "img-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- engine arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32, help="pairs per GPU per step (reference batch_size_per_gpu: 32)")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = torch.distributed.group.WORLD

    from satlas_super_resolution_b200 import _lib as L
    from satlas_super_resolution_b200 import weights
    from satlas_super_resolution_b200.ops import cur_stream
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    lib = L.load()
    B = args.batch
    cfg = dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=23), cuda_graph=not args.no_graph)
    # identical replicas on every rank (seeded init); each rank draws its own batch (manual_seed + rank, options.py:81)
    tr = ESRGANTrainer(weights.rrdbnet_state(24, 3, seed=0), weights.unet_disc_state(27, seed=1), weights.vgg19_state(seed=2),
                       cfg, device=f"cuda:{local}", process_group=pg)
    lr_h, hr_h = synthetic_batch(B, rank)
    lr_h, hr_h = lr_h.pin_memory(), hr_h.pin_memory()
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    it = [0]

    def step_resident():
        it[0] += 1
        tr._feed_kernels(tr.io, cur_stream())
        tr.optimize_parameters(it[0])

    def step_e2e():
        it[0] += 1
        tr.feed_data(lr_h, hr_h)
        tr.optimize_parameters(it[0])
        return tr.get_current_log()

    # ---- warm-up (eager first pass allocates workspaces; the next ones capture + replay the CUDA graph)
    log("trainer built")
    tr.feed_data(lr_h, hr_h)
    for i in range(args.warmup):
        step_resident()
        torch.cuda.synchronize()
        log(f"warm-up step {i} done")
    barrier()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.ssr_launch_count()
    ms_total = timed(step_resident, args.steps)
    log(f"resident timing done: {ms_total / args.steps:.2f} ms/step")
    # launches per step: graph replays do not pass through the host entry points, so count one eager step below
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    log(f"e2e timing done: {ms_e2e / args.steps:.2f} ms/step")

    # ---- one eager, instrumented step: per-launch CUDA events around the two tensor-core kernels
    tr.use_graph = False
    tr._warm.clear()
    L.check(lib.ssr_profile_start())
    l0 = lib.ssr_launch_count()
    step_resident()
    launches_per_step = lib.ssr_launch_count() - l0
    import ctypes
    ms_cls = (ctypes.c_double * 2)()
    cnt_cls = (ctypes.c_int64 * 2)()
    L.check(lib.ssr_profile_stop(ms_cls, cnt_cls, 2))
    log(f"instrumented step: conv_tc {ms_cls[0]:.2f} ms / {cnt_cls[0]} launches, wgrad_tc {ms_cls[1]:.2f} ms / {cnt_cls[1]} launches")

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_val = B * world / (ms_e2e / args.steps / 1e3)
    peak_sus, peak_burst, peak_src = measured_peaks()
    conv_ms, wgrad_ms = ms_cls[0], ms_cls[1]
    ach = FLOP_CONV_TC * B / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    ach_w = FLOP_WGRAD * B / (wgrad_ms / 1e3) / 1e12 if wgrad_ms > 0 else None
    line = {
        "metric": "img-pairs/sec 8-frame ESRGAN 4x train", "value": value, "unit": "img-pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(CONFIG, batch_per_gpu=B, global_batch=B * world, parallelism=f"dp{world}",
                       cuda_graph=not args.no_graph,
                       l2="no explicit flush: one step streams >10 GB of activations/gradients, far above the 126 MB L2"),
        "e2e": {"value": e2e_val, "unit": "img-pairs/s", "h2d_bytes_per_step": int(lr_h.numel() + hr_h.numel()),
                "d2h_bytes_per_step": 32},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "step_flop_fraction_of_peak": FLOP_STEP * B / (ms_step / 1e3) / 1e12 / peak_sus,
        "step_tflops": FLOP_STEP * B / (ms_step / 1e3) / 1e12,
        "roofline": {"kernel": "ssr::conv_chain_kernel / conv_tc_kernel (one tcgen05 implicit-GEMM body: forward + input gradient; "
                               "a dense block's five convs per chained launch)", "bound": "tensor",
                     "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus if ach else None,
                     "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches_per_step": int(cnt_cls[0]), "ms_per_step": conv_ms,
                     "flop_per_launch": FLOP_CONV_TC * B / max(1, cnt_cls[0]),
                     "traffic": _dominant_traffic().get("dram_bytes_per_launch"),
                     "traffic_note": _dominant_traffic().get("note")},
        "roofline_wgrad": {"kernel": "ssr::wgrad9_tc_kernel / wgrad_tc_kernel (tcgen05 weight gradient)", "bound": "tensor", "achieved": ach_w,
                           "peak": peak_sus, "unit": "TFLOP/s", "frac": ach_w / peak_sus if ach_w else None,
                           "launches_per_step": int(cnt_cls[1]), "ms_per_step": wgrad_ms},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        threads = usable_cores()
        log(f"cpu baseline on {threads} threads")
        t = cpu_step_time(2, 2, 1, threads)
        line["cpu_baseline"] = {"value": 2 / t, "unit": "img-pairs/s", "cores": threads, "kind": "port",
                                "sample": "2 timed optimize_parameters steps of 2 pairs (oracle/step.py, torch fp32 CPU, all host "
                                          "threads) after 1 warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
