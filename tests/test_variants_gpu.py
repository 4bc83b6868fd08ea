"""GPU parity on the other BASELINE configurations and odd shapes: 12-band input (G Cin = 96, D Cin = 99), a 48x48 low-res
tile (ragged tiles: 48 is not a multiple of the 32-wide conv tile nor a power of two for the nine-tap wgrad kernel), batch 1 / 3,
scale 8, no perceptual loss / feed_disc_lr off."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _step(cin_frames, hw, B, nb=1, percep=True, feed_disc_lr=True, seed=0):
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    d_in = 3 + (cin_frames if feed_disc_lr else 0)
    gp = nets.rrdbnet_init(cin_frames, 3, num_block=nb, seed=seed)
    dp = nets.unet_disc_init(d_in, seed=seed + 1)
    vp = losses.vgg19_init(seed=seed + 2) if percep else None
    g = torch.Generator().manual_seed(seed + 3)
    lr = torch.randint(1, 256, (B, cin_frames, hw, hw), generator=g, dtype=torch.uint8)
    hr = torch.randint(1, 256, (B, 3, 4 * hw, 4 * hw), generator=g, dtype=torch.uint8)
    opt = dict(ema_decay=0.999, lr=1e-4, perceptual=percep, feed_disc_lr=feed_disc_lr)
    orc = OracleESRGAN(gp, dp, vp, opt, num_block=nb)
    orc.feed_data(lr, hr)
    ref = orc.optimize_parameters()
    tr = ESRGANTrainer(gp, dp, vp, dict(opt, network_g=dict(num_in_ch=cin_frames, num_block=nb)))
    tr.feed_data(lr, hr)
    tr.optimize_parameters(1)
    log = tr.get_current_log()
    torch.cuda.synchronize()
    assert set(log) == set(ref)
    for k, v in ref.items():
        assert abs(log[k] - v) < 3e-2 * abs(v) + 2e-3, (k, log[k], v)
    assert rel_l2(tr.output, orc.output) < 2e-2
    # the last conv's gradient is free of ReLU-kink noise upstream of it only through d_out: a tight check on the whole backward wiring
    assert rel_l2(tr.g_grads()["conv_last.bias"], orc.g["conv_last.bias"].grad) < 5e-2
    assert rel_l2(tr.d_grads()["conv9.weight"], orc.d["conv9.weight"].grad) < 5e-2
    return tr, orc


def test_twelve_band_step():
    """BASELINE config 4: 8 frames x 12 bands = 96-channel first conv, 99-channel discriminator input"""
    _step(96, 32, 2)


def test_ragged_tile_step_48():
    _step(24, 48, 1)


def test_batch_three_no_perceptual_no_disc_lr():
    _step(24, 32, 3, percep=False, feed_disc_lr=False)


def test_scale8_forward():
    from oracle import nets
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    p = nets.rrdbnet_init(24, 3, num_block=1, scale=8, seed=4)
    x = torch.rand(1, 24, 16, 16, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = nets.rrdbnet_forward(p, x, scale=8, num_block=1)
    eng = RRDBNetEngine({k: v.cuda() for k, v in p.items()}, 24, 3, scale=8, num_block=1, want_grad=False)
    eng.repack()
    out = eng.forward(x.cuda().contiguous()).clone()
    torch.cuda.synchronize()
    assert out.shape == (1, 3, 128, 128)
    assert rel_l2(out, ref) < 1e-2
