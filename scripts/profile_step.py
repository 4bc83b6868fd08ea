"""One eager training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off ...`)."""
import sys
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import weights
from satlas_super_resolution_b200.ops import cur_stream
from satlas_super_resolution_b200.trainer import ESRGANTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tr = ESRGANTrainer(weights.rrdbnet_state(24, 3, seed=0), weights.unet_disc_state(27, seed=1), weights.vgg19_state(seed=2),
                   dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=23), cuda_graph=False))
g = torch.Generator().manual_seed(0)
lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
tr.feed_data(lr, hr)
for i in range(2):
    tr.optimize_parameters(i + 1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.feed_data(lr, hr)
tr.optimize_parameters(3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
