"""Diagnostics for tests/test_ddp_gpu.py: per-tensor relative difference between a 2-rank step and the single-process big batch.
torchrun --nproc-per-node 2 scripts/ddp_diag.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import losses, nets
from satlas_super_resolution_b200.trainer import ESRGANTrainer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gp, dp, vp = nets.rrdbnet_init(24, 3, num_block=1, seed=1), nets.unet_disc_init(27, seed=2), losses.vgg19_init(seed=3)
g = torch.Generator().manual_seed(4)
lr = torch.randint(1, 256, (4, 24, 32, 32), generator=g, dtype=torch.uint8)
hr = torch.randint(1, 256, (4, 3, 128, 128), generator=g, dtype=torch.uint8)
cfg = dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=1), cuda_graph=False)
tr = ESRGANTrainer(gp, dp, vp, cfg, device=f"cuda:{local}", process_group=dist.group.WORLD)
sl = slice(2 * rank, 2 * rank + 2)
tr.feed_data(lr[sl], hr[sl])
tr.optimize_parameters(1)
torch.cuda.synchronize()
if rank == 0:
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-20)).item()
    ones = []
    for rep in range(2):   # two single-process runs: their mutual difference is the run-to-run noise floor
        one = ESRGANTrainer(gp, dp, vp, cfg, device=f"cuda:{local}")
        one.feed_data(lr, hr)
        one.optimize_parameters(1)
        torch.cuda.synchronize()
        ones.append(({k: v.clone() for k, v in one.g_grads().items()}, {k: v.clone() for k, v in one.d_grads().items()}))
    rows = []
    for name, mine, idx in (("G", tr.g_grads(), 0), ("D", tr.d_grads(), 1)):
        for k, v in mine.items():
            rows.append((rel(v / world, ones[0][idx][k]), rel(ones[1][idx][k], ones[0][idx][k]), name + "." + k, ones[0][idx][k].norm().item()))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print(f"ddp-vs-single {r[0]:.3e}  single-vs-single {r[1]:.3e}  norm {r[3]:.3e}  {r[2]}")
dist.barrier()
dist.destroy_process_group()
