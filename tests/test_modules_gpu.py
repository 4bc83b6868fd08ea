"""GPU tests of the registry-facing surface: nn.Module forward/backward through autograd, loss modules, checkpoint
round trip, the reference-style step written with the modules vs the fused trainer step, CUDA-graph replay."""
import os
from collections import OrderedDict

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def test_generator_module_forward_backward_and_checkpoint(tmp_path):
    from oracle import nets
    from satlas_super_resolution_b200.archs import SSR_RRDBNet
    sd = nets.rrdbnet_init(24, 3, num_block=2, seed=1)
    m = SSR_RRDBNet(24, 3, num_block=2)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    x = torch.rand(2, 24, 32, 32, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = nets.rrdbnet_forward(sd, x, num_block=2)
    # eval / no_grad path (ssr/infer.py after .eval(); infer.py itself does not use no_grad -> the grad path below)
    m.eval()
    with torch.no_grad():
        y0 = m(x.cuda())
    assert rel_l2(y0, ref) < 2e-2
    # grad-recording path + backward through autograd
    m.train()
    y = m(x.cuda())
    assert y.requires_grad and torch.equal(y.detach(), y0)
    r = torch.randn_like(y) / y.numel()
    (y * r).sum().backward()
    po = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (nets.rrdbnet_forward(po, x, num_block=2) * r.cpu()).sum().backward()
    g_last = dict(m.named_parameters())["conv_last.weight"].grad
    assert rel_l2(g_last, po["conv_last.weight"].grad) < 2e-2
    g_first = dict(m.named_parameters())["conv_first.weight"].grad
    assert rel_l2(g_first, po["conv_first.weight"].grad) < 0.5       # plain oracle: LeakyReLU kink flips included
    # a second backward accumulates (torch semantics)
    y2 = m(x.cuda())
    (y2 * r).sum().backward()
    assert rel_l2(dict(m.named_parameters())["conv_last.weight"].grad, 2 * po["conv_last.weight"].grad) < 2e-2
    # torch.optim works on the parameters (views of the flat buffer) and the next forward sees the new weights
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    opt.step()
    opt.zero_grad()
    with torch.no_grad():
        y3 = m(x.cuda())
    assert not torch.equal(y3, y0)
    # checkpoint round trip in the reference's format: {'params': ..., 'params_ema': ...}
    path = os.path.join(tmp_path, "net_g.pth")
    torch.save({"params_ema": OrderedDict((k, v.cpu()) for k, v in m.state_dict().items())}, path)
    m2 = SSR_RRDBNet(24, 3, num_block=2)
    m2.load_state_dict(torch.load(path)["params_ema"], strict=True)
    m2 = m2.cuda().eval()
    with torch.no_grad():
        assert torch.equal(m2(x.cuda()), y3)


def test_discriminator_module_and_losses():
    from oracle import losses as olosses
    from oracle import nets
    from satlas_super_resolution_b200.archs import SSR_UNetDiscriminatorSN
    from satlas_super_resolution_b200.losses import GANLoss, L1Loss, PerceptualLoss
    sd = nets.unet_disc_init(27, seed=3)
    d = SSR_UNetDiscriminatorSN(27)
    d.load_state_dict(sd, strict=True)
    d = d.cuda().train()
    x = torch.rand(2, 27, 128, 128, generator=torch.Generator().manual_seed(4))
    xg = x.cuda().requires_grad_(True)
    pred = d(xg)
    po = {k: (v.clone().requires_grad_(True) if not k.endswith(("weight_u", "weight_v")) else v.clone()) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    pred_ref = nets.unet_disc_forward(po, xo, training=True)
    assert rel_l2(pred, pred_ref) < 2e-2
    gan = GANLoss("vanilla", loss_weight=0.1)
    l = gan(pred, True, is_disc=False)
    l_ref = olosses.gan_loss_vanilla(pred_ref, True, is_disc=False, loss_weight=0.1)
    assert abs(l.item() - l_ref.item()) < 2e-3 * abs(l_ref.item()) + 1e-5
    l.backward()
    l_ref.backward()
    assert rel_l2(xg.grad, xo.grad) < 0.2
    assert rel_l2(dict(d.named_parameters())["conv9.weight"].grad, po["conv9.weight"].grad) < 2e-2
    assert torch.allclose(d.state_dict()["conv2.weight_u"].cpu(), po["conv2.weight_u"], atol=1e-5)   # u advanced once
    # frozen discriminator (generator step): no parameter grads, input grad still flows
    for p in d.parameters():
        p.requires_grad = False
        p.grad = None
    xg2 = x.cuda().requires_grad_(True)
    gan(d(xg2), True).backward()
    assert xg2.grad is not None and all(p.grad is None for p in d.parameters())
    # L1 and perceptual loss modules
    a = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(5))
    b = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(6))
    ag = a.cuda().requires_grad_(True)
    l1 = L1Loss(loss_weight=1.0)(ag, b.cuda())
    assert abs(l1.item() - (a - b).abs().mean().item()) < 1e-5
    l1.backward()
    assert torch.allclose(ag.grad.cpu(), torch.sign(a - b) / a.numel(), atol=1e-9)
    per = PerceptualLoss(olosses.DEFAULT_LAYER_WEIGHTS, vgg_seed=7)
    ag2 = a.cuda().requires_grad_(True)
    lp, style = per(ag2, b.cuda())
    assert style is None
    ref = olosses.perceptual_loss(olosses.vgg19_init(seed=7), a, b)
    # weights.vgg19_state and oracle.losses.vgg19_init draw the same seeded tensors
    assert abs(lp.item() - ref.item()) < 2e-2 * ref.item()
    lp.backward()
    assert ag2.grad is not None and ag2.grad.abs().sum() > 0


def _opt(tmp_path, nb=2):
    return {
        "name": "t", "model_type": "SSRESRGANModel", "scale": 4, "num_gpu": 1, "is_train": True, "dist": False,
        "l1_gt_usm": True, "percep_gt_usm": True, "gan_gt_usm": False, "feed_disc_lr": True, "cuda_graph": False,
        "network_g": dict(type="SSR_RRDBNet", num_in_ch=24, num_out_ch=3, num_feat=64, num_block=nb, num_grow_ch=32),
        "network_d": dict(type="SSR_UNetDiscriminatorSN", num_in_ch=27, num_feat=64, skip_connection=True),
        "path": {"experiments_root": str(tmp_path), "models": str(tmp_path / "models"),
                 "training_states": str(tmp_path / "training_states")},
        "train": {"ema_decay": 0.999,
                  "optim_g": dict(type="Adam", lr=1e-4, weight_decay=0, betas=[0.9, 0.99]),
                  "optim_d": dict(type="Adam", lr=1e-4, weight_decay=0, betas=[0.9, 0.99]),
                  "scheduler": dict(type="MultiStepLR", milestones=[3], gamma=0.5),
                  "pixel_opt": dict(type="L1Loss", loss_weight=1.0, reduction="mean"),
                  "perceptual_opt": dict(type="PerceptualLoss", layer_weights={"conv1_2": 0.1, "conv2_2": 0.1, "conv3_4": 1,
                                                                                "conv4_4": 1, "conv5_4": 1},
                                         vgg_type="vgg19", use_input_norm=True, perceptual_weight=1.0, style_weight=0,
                                         range_norm=False, criterion="l1", vgg_seed=0),
                  "gan_opt": dict(type="GANLoss", gan_type="vanilla", real_label_val=1.0, fake_label_val=0.0, loss_weight=0.1),
                  "net_d_iters": 1, "net_d_init_iters": 0},
    }


def _batch(B=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
    hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    return {"lr": lr, "hr": hr}


def test_model_from_opt_matches_reference_style_step(tmp_path):
    """SSRESRGANModel.optimize_parameters (fused trainer) vs the reference's own statement sequence
    (ssr_esrgan_model.py:119-233) executed with OUR registry modules + torch autograd + torch.optim.Adam."""
    from satlas_super_resolution_b200.losses import GANLoss, L1Loss, PerceptualLoss
    from satlas_super_resolution_b200.registry import build_model, build_network
    opt = _opt(tmp_path)
    torch.manual_seed(0)
    model = build_model(opt)
    data = _batch()
    # ---- module-level replica starting from the same weights
    net_g = build_network(opt["network_g"])
    net_g.load_state_dict(model.net_g.state_dict())
    net_d = build_network(opt["network_d"])
    net_d.load_state_dict(model.net_d.state_dict())
    net_g, net_d = net_g.cuda().train(), net_d.cuda().train()
    cri_pix, cri_gan = L1Loss(1.0), GANLoss("vanilla", loss_weight=0.1)
    cri_per = PerceptualLoss(opt["train"]["perceptual_opt"]["layer_weights"], vgg_seed=0)
    opt_g = torch.optim.Adam(net_g.parameters(), lr=1e-4, betas=(0.9, 0.99))
    opt_d = torch.optim.Adam(net_d.parameters(), lr=1e-4, betas=(0.9, 0.99))

    model.feed_data(data)
    lr, gt, gt_usm = model.lr.clone(), model.gt.clone(), model.gt_usm.clone()
    model.optimize_parameters(1)
    log = model.get_current_log()

    lr_resized = F.interpolate(lr, scale_factor=4)
    for p in net_d.parameters():
        p.requires_grad = False
    opt_g.zero_grad()
    output = net_g(lr)
    l_pix = cri_pix(output, gt_usm)
    l_per, _ = cri_per(output, gt_usm)
    fake_g_pred = net_d(torch.cat((output, lr_resized), 1))
    l_gan = cri_gan(fake_g_pred, True, is_disc=False)
    (l_pix + l_per + l_gan).backward()
    opt_g.step()
    for p in net_d.parameters():
        p.requires_grad = True
    opt_d.zero_grad()
    real_pred = net_d(torch.cat((gt, lr_resized), 1))
    l_d_real = cri_gan(real_pred, True, is_disc=True)
    l_d_real.backward()
    fake_pred = net_d(torch.cat((output, lr_resized), 1).detach().clone())
    l_d_fake = cri_gan(fake_pred, False, is_disc=True)
    l_d_fake.backward()
    opt_d.step()
    torch.cuda.synchronize()
    ref = dict(l_g_pix=l_pix.item(), l_g_percep=l_per.item(), l_g_gan=l_gan.item(), l_d_real=l_d_real.item(),
               out_d_real=real_pred.mean().item(), l_d_fake=l_d_fake.item(), out_d_fake=fake_pred.mean().item())
    for k, v in ref.items():
        assert abs(log[k] - v) < 1e-3 * abs(v) + 1e-4, (k, log[k], v)
    assert rel_l2(model.output, output) < 1e-3
    # both paths moved the weights the same way (Adam's first step is ~lr*sign(g): elements whose gradient is at the 1e-8
    # eps level are decided by atomic-add ordering noise, hence a relative tolerance of a few % of the 1e-4 step)
    for k in ("conv_first.weight", "body.1.rdb2.conv3.weight", "conv_last.bias"):
        assert rel_l2(model.net_g.state_dict()[k], net_g.state_dict()[k]) < 1e-3, k
    for k in ("conv0.weight", "conv5.weight_orig", "conv5.weight_u"):
        assert rel_l2(model.net_d.state_dict()[k], net_d.state_dict()[k]) < 1e-3, k


def test_model_lr_schedule_save_resume_and_test(tmp_path):
    from satlas_super_resolution_b200.registry import build_model
    opt = _opt(tmp_path, nb=1)
    model = build_model(opt)
    data = _batch(seed=1)
    lrs = []
    for it in range(1, 6):
        model.update_learning_rate(it, warmup_iter=-1)
        model.feed_data(data)
        model.optimize_parameters(it)
        lrs.append(model.get_current_learning_rate()[0])
    assert lrs[:3] == [1e-4, 1e-4, 1e-4] and abs(lrs[4] - 5e-5) < 1e-12      # MultiStepLR milestone 3, gamma 0.5
    log = model.get_current_log()
    assert set(log) == {"l_g_pix", "l_g_percep", "l_g_gan", "l_d_real", "out_d_real", "l_d_fake", "out_d_fake"}
    model.save(0, 5)
    g = torch.load(tmp_path / "models" / "net_g_5.pth")
    assert set(g) == {"params", "params_ema"} and len(g["params"]) == len(model.net_g.state_dict())
    assert set(torch.load(tmp_path / "models" / "net_d_5.pth")) == {"params"}
    state = torch.load(tmp_path / "training_states" / "5.state")
    assert state["iter"] == 5 and len(state["optimizers"]) == 2
    # EMA lags the raw weights; test() runs the EMA generator in eval mode
    assert not torch.equal(g["params"]["conv_first.weight"], g["params_ema"]["conv_first.weight"])
    model.test()
    assert model.output.shape == (2, 3, 128, 128)
    vis = model.get_current_visuals()
    assert set(vis) == {"lr", "result", "gt"}
    # resume into a fresh model
    opt2 = _opt(tmp_path, nb=1)
    opt2["path"]["pretrain_network_g"] = str(tmp_path / "models" / "net_g_5.pth")
    opt2["path"]["pretrain_network_d"] = str(tmp_path / "models" / "net_d_5.pth")
    m2 = build_model(opt2)
    m2.resume_training(state)
    assert torch.equal(m2.net_g.state_dict()["conv_body.weight"].cpu(), g["params"]["conv_body.weight"])
    assert torch.equal(m2.net_g_ema.state_dict()["conv_body.weight"].cpu(), g["params_ema"]["conv_body.weight"])
    assert m2.trainer.opt_g.step_count == 5


def test_test_only_model_builds_generator_only(tmp_path):
    """basicsr SRModel with is_train False (ssr/test.py): net_g only -- no network_d / losses / optimizers needed in the option file"""
    from oracle import nets
    from satlas_super_resolution_b200.registry import build_model
    p = nets.rrdbnet_init(24, 3, num_feat=32, num_block=1, num_grow_ch=16, seed=9)
    torch.save({"params_ema": p}, tmp_path / "g.pth")
    opt = {"name": "t", "model_type": "SSRESRGANModel", "scale": 4, "num_gpu": 1, "is_train": False, "dist": False,
           "network_g": dict(type="SSR_RRDBNet", num_in_ch=24, num_out_ch=3, num_feat=32, num_block=1, num_grow_ch=16),
           "path": {"pretrain_network_g": str(tmp_path / "g.pth"), "param_key_g": "params_ema", "strict_load_g": True}}
    model = build_model(opt)
    assert model.trainer is None and not hasattr(model, "net_d") and model.optimizers == []
    data = _batch(seed=4)
    model.feed_data(data)
    model.test()
    with torch.no_grad():
        ref = nets.rrdbnet_forward(p, data["lr"].float() / 255, num_block=1)
    assert rel_l2(model.output, ref) < 1e-2
    assert set(model.get_current_visuals()) == {"lr", "result", "gt"}


def test_discriminator_backward_of_a_stale_forward_raises():
    """two forwards then one backward over both graphs would use the second forward's spectral-norm state for the first: refuse"""
    from satlas_super_resolution_b200.archs import SSR_UNetDiscriminatorSN
    torch.manual_seed(0)
    d = SSR_UNetDiscriminatorSN(3).cuda().train()
    x = torch.rand(1, 3, 64, 64, device="cuda")
    a = d(x).mean()
    b = d(x * 0.5).mean()
    with pytest.raises(RuntimeError, match="no longer the latest"):
        (a + b).backward()


def test_cuda_graph_replay_equals_eager():
    from oracle import losses, nets
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    gp, dp, vp = nets.rrdbnet_init(24, 3, num_block=1, seed=1), nets.unet_disc_init(27, seed=2), losses.vgg19_init(seed=3)
    data = _batch(seed=2)
    outs = []
    for graph in (False, True, False):
        tr = ESRGANTrainer(gp, dp, vp, dict(ema_decay=0.999, lr=1e-4, network_g=dict(num_in_ch=24, num_block=1), cuda_graph=graph))
        grads2 = None
        for it in range(1, 5):
            tr.feed_data(data["lr"], data["hr"])
            tr.optimize_parameters(it)           # it=1 eager (both), it=2 capture+replay (graph run), it>=3 replays
            if it == 2:
                torch.cuda.synchronize()
                grads2 = {k: v.clone() for k, v in tr.g_grads().items()}
        torch.cuda.synchronize()
        outs.append((tr.g_state_dict(), tr.d_state_dict(), tr.get_current_log(), tr.opt_g.step_count, grads2,
                     tr.opt_g.hyper_dev.cpu() if graph else None))
    assert outs[0][3] == outs[1][3] == 4
    # the device-side step counter advanced with the host one: t = 4, 1 - 0.9^4, sqrt(1 - 0.99^4)
    h = outs[1][5]
    assert h[3].item() == 4 and abs(h[1].item() - (1 - 0.9 ** 4)) < 1e-6 and abs(h[2].item() - (1 - 0.99 ** 4) ** 0.5) < 1e-6
    # Two EAGER runs already differ from each other after a step (f32 atomic-add ordering -> Adam's sign-like first
    # update at eps-level gradients -> ReLU / L1 kink flips in the next step); the graph run must sit inside that same
    # run-to-run spread: outs[0], outs[2] = eager twice, outs[1] = graph.
    for k in ("conv_first.weight", "body.0.rdb3.conv5.weight", "conv_last.bias"):
        spread = rel_l2(outs[2][4][k], outs[0][4][k])
        dev = rel_l2(outs[1][4][k], outs[0][4][k])
        print(f"  step-2 grad {k}: eager-vs-eager {spread:.3e}  graph-vs-eager {dev:.3e}")
        # the run-to-run spread is itself one draw of a heavy-tailed quantity: over six repetitions on the B200 the eager-vs-eager
        # difference of the 3-element conv_last.bias ranged 2.7e-3 .. 3.0e-2 while graph-vs-eager stayed at 6e-3 .. 1.2e-2
        # (profiles/r02b_logs/t_graph_repeats.log): a floor of 5e-2 keeps a low draw of `spread` from failing a correct replay;
        # a replay that skipped or duplicated work shows up as O(1)
        assert dev < 3 * spread + 5e-2, k
    for k in ("conv_first.weight", "body.0.rdb3.conv5.weight"):
        spread = rel_l2(outs[2][0][k], outs[0][0][k])
        assert rel_l2(outs[1][0][k], outs[0][0][k]) < 3 * spread + 1e-4, k
    assert rel_l2(outs[1][1]["conv4.weight_orig"], outs[0][1]["conv4.weight_orig"]) < 1e-3
    for k, v in outs[0][2].items():
        assert abs(outs[1][2][k] - v) < 2e-3 * abs(v) + 2e-4
