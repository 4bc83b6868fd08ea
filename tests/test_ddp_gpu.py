"""2-GPU data-parallel equivalence (runs only where >= 2 devices are visible, e.g. `gpurun --gpus 2`): two ranks, each with half
of a batch, must produce the same averaged gradients and losses as one process fed the whole batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
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
log = tr.get_current_log()
torch.cuda.synchronize()
if rank == 0:
    one = ESRGANTrainer(gp, dp, vp, cfg, device=f"cuda:{local}")
    one.feed_data(lr, hr)
    one.optimize_parameters(1)
    ref = one.get_current_log()
    torch.cuda.synchronize()
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-20)).item()
    worst = 0.0
    for k, v in tr.g_grads().items():
        worst = max(worst, rel(v / world, one.g_grads()[k]))
    for k, v in tr.d_grads().items():
        worst = max(worst, rel(v / world, one.d_grads()[k]))
    for k in ref:
        assert abs(log[k] - ref[k]) < 1e-3 * abs(ref[k]) + 1e-4, (k, log[k], ref[k])
    assert worst < 2e-3, worst
    for k in ("conv_first.weight", "conv_last.weight"):
        assert rel(tr.g_state_dict()[k], one.g_state_dict()[k]) < 1e-3
    print("ddp-ok worst grad rel", worst)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_process(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ddp-ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
