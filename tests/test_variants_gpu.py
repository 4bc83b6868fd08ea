"""GPU parity on the other BASELINE configurations and odd shapes: 12-band input (G Cin = 96, D Cin = 99), a 48x48 low-res
tile (ragged tiles: 48 is not a multiple of the 32-wide conv tile nor a power of two for the nine-tap wgrad kernel), batch 1 / 3,
scale 8, no perceptual loss / feed_disc_lr off."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _step(cin_frames, hw, B, nb=1, percep=True, feed_disc_lr=True, seed=0, old_hr=False, extra_opt=None, iters=(1,)):
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    d_in = 3 + (cin_frames if feed_disc_lr else 0) + (3 if old_hr else 0)
    gp = nets.rrdbnet_init(cin_frames, 3, num_block=nb, seed=seed)
    dp = nets.unet_disc_init(d_in, seed=seed + 1)
    vp = losses.vgg19_init(seed=seed + 2) if percep else None
    g = torch.Generator().manual_seed(seed + 3)
    lr = torch.randint(1, 256, (B, cin_frames, hw, hw), generator=g, dtype=torch.uint8)
    hr = torch.randint(1, 256, (B, 3, 4 * hw, 4 * hw), generator=g, dtype=torch.uint8)
    old = torch.randint(1, 256, (B, 3, 4 * hw, 4 * hw), generator=g, dtype=torch.uint8) if old_hr else None
    opt = dict(ema_decay=0.999, lr=1e-4, perceptual=percep, feed_disc_lr=feed_disc_lr, **(extra_opt or {}))
    orc = OracleESRGAN(gp, dp, vp, opt, num_block=nb)
    tr = ESRGANTrainer(gp, dp, vp, dict(opt, network_g=dict(num_in_ch=cin_frames, num_block=nb)))
    for it in iters:
        orc.feed_data(lr, hr, old)
        ref = orc.optimize_parameters(it)
        tr.feed_data(lr, hr, old)
        tr.optimize_parameters(it)
        log = tr.get_current_log()
        torch.cuda.synchronize()
        assert set(log) == set(ref), (it, set(log), set(ref))
        for k, v in ref.items():
            assert abs(log[k] - v) < 3e-2 * abs(v) + 2e-3, (it, k, log[k], v)
        assert rel_l2(tr.output, orc.output) < 2e-2
    if "l_g_pix" not in ref:
        return tr, orc
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


def test_old_hr_discriminator_conditioning():
    """ssr_esrgan_model.py:112-114, 171-174, 202-207: the discriminator sees [image | lr_resized | old_hr] (30 channels)"""
    tr, orc = _step(24, 32, 2, old_hr=True)
    assert tr.d_in_ch == 30
    assert rel_l2(tr.d_grads()["conv0.weight"], orc.d["conv0.weight"].grad) < 0.2
    # without the low-res stack: [image | old_hr] (6 channels)
    _step(24, 32, 1, feed_disc_lr=False, old_hr=True, seed=3)


def test_net_d_iters_runs_ema_every_iteration_and_separate_optim_d():
    """net_d_iters = 2: iteration 1 has no generator step (no l_g_* in the log) but model_ema still runs (:230-231);
    train.optim_d carries its own lr / betas (esrgan_s2naip_urban.yml:103-107)."""
    tr, orc = _step(24, 32, 2, extra_opt=dict(net_d_iters=2, lr_d=3e-4, betas_d=(0.5, 0.9)), iters=(1, 2))
    g_ema = tr.g_state_dict(ema=True)
    for k in ("conv_first.weight", "body.0.rdb2.conv3.weight", "conv_last.bias"):
        assert rel_l2(g_ema[k], orc.g_ema[k]) < 1e-5, k
    # D moved with ITS learning rate: after two steps |delta| is ~2 * 3e-4 per element, three times what optim_g's lr would give
    d1, d0 = tr.d_state_dict(), orc.d
    step = (d1["conv9.weight"].cpu() - nets_init_d(27)["conv9.weight"]).abs().mean().item()
    assert 3e-4 < step < 7e-4, step
    # (Adam's sign-like step on elements whose gradient sits at the 1e-8 eps level is decided by rounding: a few 1e-3 of the weights)
    assert rel_l2(d1["conv9.weight"], d0["conv9.weight"]) < 1e-2


def nets_init_d(cin, seed=1):
    from oracle import nets
    return nets.unet_disc_init(cin, seed=seed)


@pytest.mark.parametrize("scale", [2, 1])
def test_pixel_unshuffle_front_end(scale):
    """rrdbnet_arch.py:95-98, 117-120: scale 2 / 1 put pixel_unshuffle(x, 2 / 4) in front of conv_first (fused into the ingest)"""
    from oracle import nets
    from satlas_super_resolution_b200.archs import SSR_RRDBNet
    p = nets.rrdbnet_init(3, 3, num_block=1, scale=scale, seed=6)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        ref = nets.rrdbnet_forward(p, x, scale=scale, num_block=1)
    net = SSR_RRDBNet(3, 3, scale=scale, num_block=1)
    net.load_state_dict(p, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda())
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (2, 3, 64 * scale, 64 * scale)
    assert rel_l2(out, ref) < 1e-2
    # training mode: gradients flow to conv_first through the unshuffled input
    net.train()
    out = net(x.cuda())
    out.mean().backward()
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    nets.rrdbnet_forward(po, x, scale=scale, num_block=1).mean().backward()
    # d_out = 1 / numel is not a bf16 number: the back-propagated constant carries its 0.2 % rounding
    assert rel_l2(net.conv_last.bias.grad, po["conv_last.bias"].grad) < 5e-3
    assert rel_l2(net.conv_first.weight.grad, po["conv_first.weight"].grad) < 0.3


def test_ssim_loss_kernel_vs_oracle():
    """SSIMLoss (ssr/losses/basic_loss.py:50-60): loss value and dLoss/dx of ssr_ssim_loss against autograd of the oracle's torch
    restatement of kornia.losses.ssim_loss -- odd sizes, so every reflected border row / column of the adjoint filter is hit"""
    from oracle import losses
    from satlas_super_resolution_b200.losses import SSIMLoss
    g = torch.Generator().manual_seed(11)
    for shape, w in (((2, 3, 128, 128), 1.0), ((3, 2, 17, 9), 0.37), ((1, 1, 3, 5), 2.0)):
        x = torch.rand(shape, generator=g)
        y = (x + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
        xo = x.clone().requires_grad_(True)
        ref = losses.ssim_loss(xo, y, w)
        ref.backward()
        xe = x.cuda().requires_grad_(True)
        out = SSIMLoss(loss_weight=w)(xe, y.cuda())
        out.backward()
        torch.cuda.synchronize()
        assert abs(out.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), (shape, out.item(), ref.item())
        assert rel_l2(xe.grad, xo.grad) < 1e-4, (shape, rel_l2(xe.grad, xo.grad))
    # identical images: ssim == 1 everywhere, loss 0, zero gradient
    xe = x.cuda().requires_grad_(True)
    out = SSIMLoss()(xe, x.cuda())
    out.backward()
    assert abs(out.item()) < 1e-6 and xe.grad.abs().max().item() < 1e-6


def test_step_with_ssim_opt():
    """train.ssim_opt (ssr_esrgan_model.py:87-90, 163-166): l_g_ssim joins the generator loss and the log"""
    tr, orc = _step(24, 32, 2, extra_opt=dict(ssim_weight=0.5))
    assert "l_g_ssim" in tr.get_current_log()
    assert list(tr.get_current_log())[:4] == ["l_g_pix", "l_g_percep", "l_g_ssim", "l_g_gan"]
