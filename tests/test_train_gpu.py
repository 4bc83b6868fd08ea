"""GPU parity of the training-side engines against the fp32 CPU oracle (autograd over oracle/nets.py, oracle/losses.py):
discriminator forward / input gradient / weight gradients (spectral norm included), the VGG perceptual loss and its
image gradient, the generator backward, and one full optimize_parameters step.

Tolerances (bf16 operands, fp32 accumulate; see DESIGN.md "precision contract"):
  forward tensors and losses: relative L2 <= 2e-2;  gradients: relative L2 <= 5e-2 per tensor (cosine >= 0.998).
"""
import ctypes as C
from collections import OrderedDict

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-20)).item()


def cosine(got, ref):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    return (torch.dot(got, ref) / (got.norm() * ref.norm() + 1e-30)).item()


FAILS = []


def report(tag, got, ref, tol):
    e, c = rel_l2(got, ref), cosine(got, ref)
    flag = "" if e < tol else "   <-- over tolerance"
    print(f"  {tag:40s} rel_l2={e:.3e} cos={c:.6f}{flag}")
    if not e < tol:
        FAILS.append(f"{tag}: rel_l2 {e:.3e} >= {tol}")
    return e


def check_fails():
    msgs = list(FAILS)
    FAILS.clear()
    assert not msgs, "; ".join(msgs)


def nchw_mask(act_t, lo=0, hi=None):
    """engine NHWC bf16 activation -> bool NCHW cpu mask of its positive entries"""
    t = act_t[..., lo:hi] if hi is not None else act_t[..., lo:]
    return (t > 0).permute(0, 3, 1, 2).cpu()


def test_discriminator_forward_backward():
    """Forward vs the plain oracle; backward vs the oracle evaluated WITH THE ENGINE'S ACTIVATION PATTERN (the LeakyReLU
    masks the engine saved), which removes kink flips caused by bf16 forward noise and leaves only rounding."""
    from oracle import nets
    from satlas_super_resolution_b200 import _lib as L
    from satlas_super_resolution_b200.discriminator import UNetDiscEngine
    from satlas_super_resolution_b200.ops import lib
    B, cin, H = 2, 27, 128
    p = nets.unet_disc_init(cin, seed=1)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, cin, H, H, generator=g)
    xb = x.to(torch.bfloat16).float().requires_grad_(True)
    mk = lambda: {k: (v.clone().requires_grad_(True) if not k.endswith(("weight_u", "weight_v")) else v.clone())
                  for k, v in p.items()}
    po = mk()
    with torch.no_grad():
        logits_ref = nets.unet_disc_forward(po, xb, training=True)
    dl = torch.randn(B, 1, H, H, generator=g) / (B * H * H)

    pc = {k: v.cuda().contiguous() for k, v in p.items()}
    grads = {k: torch.zeros_like(v) for k, v in pc.items() if not k.endswith(("weight_u", "weight_v"))}
    eng = UNetDiscEngine(pc, cin, grads=grads)
    ws = eng.workspace(B, H, H)
    L.check(lib().ssr_ingest_nchw(x.cuda().data_ptr(), L.SSR_F32, ws.x_in.ptr(), ws.x_in.stride, B, cin, H, H, 32, 1.0, None,
                                  None, None))
    logits = eng.forward(ws, training=True)
    torch.cuda.synchronize()
    print()
    report("logits", logits, logits_ref, 2e-2)
    for name in ("conv1", "conv4", "conv8"):
        report(f"{name}.weight_u", pc[f"{name}.weight_u"], po[f"{name}.weight_u"], 1e-3)
        report(f"{name}.weight_v", pc[f"{name}.weight_v"], po[f"{name}.weight_v"], 1e-3)
    # oracle gradients with the engine's activation pattern
    acts = dict(conv0=ws.x0, conv1=ws.x1, conv2=ws.x2, conv3=ws.x3, conv4=ws.a4, conv5=ws.a5, conv6=ws.a6, conv7=ws.a7,
                conv8=ws.a8)
    masks = {k: nchw_mask(a.t) for k, a in acts.items()}
    pm = mk()
    out_m = nets.unet_disc_forward(pm, xb, training=True, act=nets.masked_lrelu(masks))
    report("pattern-forced oracle vs plain oracle", out_m, logits_ref, 1e-2)
    out_m.backward(dl)
    eng.backward(ws, dl.cuda().contiguous(), need_wgrad=True, need_dinput=True)
    torch.cuda.synchronize()
    d_in = ws.d_in.t[..., :cin].permute(0, 3, 1, 2).float()
    report("d_input", d_in, xb.grad, 2e-2)
    for k, gr in grads.items():
        report(f"grad {k}", gr, pm[k].grad, 2e-2)
    # eval mode: no power iteration, sigma from the stored u, v
    u_before = pc["conv3.weight_u"].clone()
    logits_eval = eng.forward(ws, training=False).clone()
    torch.cuda.synchronize()
    assert torch.equal(u_before, pc["conv3.weight_u"])
    with torch.no_grad():
        ref_eval = nets.unet_disc_forward(po, xb, training=False)
    report("logits (eval mode)", logits_eval, ref_eval, 2e-2)
    check_fails()


def test_perceptual_loss_and_gradient():
    from oracle import losses
    from satlas_super_resolution_b200.vgg import PerceptualEngine
    B, H = 2, 128
    vp = losses.vgg19_init(seed=3)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(B, 3, H, H, generator=g).requires_grad_(True)
    gt = torch.rand(B, 3, H, H, generator=g)
    with torch.no_grad():
        ref = losses.perceptual_loss(vp, x, gt)
    lw = losses.DEFAULT_LAYER_WEIGHTS
    eng = PerceptualEngine({k: v.cuda() for k, v in vp.items()}, lw)
    loss = torch.zeros(1, device="cuda")
    dx = torch.zeros(B, 3, H, H, device="cuda")
    eng.loss_and_grad(x.detach().cuda().contiguous(), gt.cuda().contiguous(), loss, dx)
    torch.cuda.synchronize()
    print()
    print(f"  perceptual loss: engine {loss.item():.6f} oracle {ref.item():.6f}")
    assert abs(loss.item() - ref.item()) / ref.item() < 2e-2
    # gradient: oracle evaluated with the engine's ReLU pattern, pool arg-max choices and L1 signs
    ws = eng.workspace(B, H, H)
    relu_masks, pool_index, signs = {}, {}, {}
    last_feat = None
    for item in ws.order:
        if item[0] == "conv":
            _, name, _src, out, hh, ww, _c = item
            full = out.t.float().permute(0, 3, 1, 2).cpu()          # 2B images
            relu_masks[name] = full[:B] > 0
            if name in lw:
                signs[name] = torch.sign(full[:B] - full[B:])
                last_feat = full[:B]
        else:
            _, pname, _src, _dst, hh, ww = item
            f = torch.relu(last_feat)
            n, c, h, w = f.shape
            win = f.view(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
            pool_index[pname] = win.argmax(dim=-1)
    relu_fn, pool_fn = losses.masked_relu_pool(relu_masks, pool_index)
    fx = losses.vgg19_features(vp, x, lw.keys(), relu_fn=relu_fn, pool_fn=pool_fn)
    with torch.no_grad():
        fg = losses.vgg19_features(vp, gt, lw.keys())
    total = sum((signs[k] * (fx[k] - fg[k])).mean() * w for k, w in lw.items())
    total.backward()
    report("d perceptual / d x", dx, x.grad, 3e-2)
    check_fails()


def _g_setup(num_block, B, seed=5):
    from oracle import nets
    p = nets.rrdbnet_init(24, 3, num_block=num_block, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 24, 32, 32, generator=g)
    d_out = torch.randn(B, 3, 128, 128, generator=g) / (B * 3 * 128 * 128)
    return p, x, d_out


@pytest.mark.parametrize("num_block,tmem", [(1, False), (3, False), (3, True)])
def test_generator_backward(num_block, tmem, monkeypatch):
    """tmem: the dense-block input-gradient chain keeps its running sum in tensor memory (SSR_DGRAD_TMEM=1)"""
    from oracle import nets
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    monkeypatch.setenv("SSR_DGRAD_TMEM", "1" if tmem else "0")
    B = 2
    p, x, d_out = _g_setup(num_block, B)
    with torch.no_grad():
        out_ref = nets.rrdbnet_forward(p, x, num_block=num_block)
    pc = {k: v.cuda().contiguous() for k, v in p.items()}
    grads = {k: torch.zeros_like(v) for k, v in pc.items()}
    eng = RRDBNetEngine(pc, 24, 3, num_block=num_block, want_grad=True, grads=grads)
    assert eng.dgrad_tmem == tmem
    eng.repack()
    out = eng.forward(x.cuda().contiguous(), train=True)
    eng.backward(d_out.cuda().contiguous(), B, 32, 32)
    torch.cuda.synchronize()
    print()
    report("output", out, out_ref, 2e-2)
    # oracle with the engine's LeakyReLU pattern
    ws = eng.workspace(B, 32, 32, True)
    masks = {}
    for i in range(num_block):
        for j in range(3):
            buf = ws.bufs[3 * i + j].t
            for k in range(1, 5):
                lo = 64 + 32 * (k - 1)
                masks[f"body.{i}.rdb{j + 1}.conv{k}"] = nchw_mask(buf, lo, lo + 32)
    masks["conv_up1"], masks["conv_up2"], masks["conv_hr"] = nchw_mask(ws.up_out[0].t), nchw_mask(ws.up_out[1].t), nchw_mask(ws.hr.t)
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out_m = nets.rrdbnet_forward(po, x, num_block=num_block, act=nets.masked_lrelu(masks))
    report("pattern-forced oracle vs plain oracle", out_m, out_ref, 1e-2)
    out_m.backward(d_out)
    worst = 0.0
    for k in grads:
        e = rel_l2(grads[k], po[k].grad)
        worst = max(worst, e)
        if e > 2e-2 or k in ("conv_first.weight", "conv_last.weight", "conv_last.bias", "body.0.rdb1.conv1.weight",
                             "body.0.rdb3.conv5.weight", "body.0.rdb2.conv3.bias", "conv_up1.weight", "conv_body.weight"):
            print(f"  grad {k:34s} rel_l2={e:.3e} cos={cosine(grads[k], po[k].grad):.6f}")
    print(f"  worst gradient rel_l2 over {len(grads)} tensors: {worst:.3e}")
    check_fails()
    assert worst < 2e-2


def test_full_step_vs_oracle():
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    B, nb = 2, 2
    gp = nets.rrdbnet_init(24, 3, num_block=nb, seed=7)
    dp = nets.unet_disc_init(27, seed=8)
    vp = losses.vgg19_init(seed=9)
    g = torch.Generator().manual_seed(10)
    lr_u8 = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
    hr_u8 = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    # smooth the HR tile a little so the USM mask is not saturated everywhere
    hr_u8 = F.avg_pool2d(hr_u8.float(), 5, 1, 2).round().clamp(0, 255).to(torch.uint8)
    opt = dict(ema_decay=0.999, lr=1e-4)
    orc = OracleESRGAN(gp, dp, vp, opt, num_block=nb)
    orc.feed_data(lr_u8, hr_u8)
    ref_log = orc.optimize_parameters()

    tr = ESRGANTrainer(gp, dp, vp, dict(opt, network_g=dict(num_in_ch=24, num_block=nb)))
    tr.feed_data(lr_u8.pin_memory(), hr_u8.pin_memory())
    torch.cuda.synchronize()
    print()
    report("gt_usm (USM sharpen)", tr.gt_usm, orc.gt_usm, 1e-3)
    g0 = tr.g_state_dict()
    d0 = tr.d_state_dict()
    tr.optimize_parameters(1)
    log = tr.get_current_log()
    torch.cuda.synchronize()
    for k in ref_log:
        print(f"  {k:12s} engine {log[k]:+.6f} oracle {ref_log[k]:+.6f}")
        tol = 3e-2 * abs(ref_log[k]) + 2e-3
        assert abs(log[k] - ref_log[k]) < tol, k
    report("output", tr.output, orc.output, 2e-2)
    # gradients left in the flat buffers
    # (plain oracle here: ReLU / L1 kinks flipped by bf16 forward noise dominate this comparison, so only a coarse
    # bound is asserted; the exact backward check is the pattern-forced tests above)
    worst_g = max(rel_l2(v, orc.g[k].grad) for k, v in tr.g_grads().items())
    worst_d = max(rel_l2(v, orc.d[k].grad) for k, v in tr.d_grads().items())
    print(f"  worst G grad rel_l2 {worst_g:.3e}   worst D grad rel_l2 {worst_d:.3e}")
    assert worst_g < 0.5 and worst_d < 0.2
    # Adam's first step moves every weight by ~lr*sign(grad): compare the UPDATE direction
    g1, d1 = tr.g_state_dict(), tr.d_state_dict()
    upd_cos = []
    for k in ("conv_first.weight", "body.0.rdb1.conv1.weight", "body.1.rdb3.conv5.weight", "conv_last.weight"):
        du = (g1[k] - g0[k]).cpu()
        dr = (orc.g[k].detach() - gp[k])
        upd_cos.append(cosine(du, dr))
        print(f"  update cos {k:30s} {upd_cos[-1]:.4f}")
    assert min(upd_cos) > 0.9
    for k in ("conv0.weight", "conv3.weight_orig", "conv9.weight"):
        du = (d1[k] - d0[k]).cpu()
        dr = (orc.d[k].detach() - dp[k])
        c = cosine(du, dr)
        print(f"  update cos D {k:28s} {c:.4f}")
        assert c > 0.9
    # EMA and spectral-norm state
    report("ema conv_first.weight", tr.g_state_dict(ema=True)["conv_first.weight"], orc.g_ema["conv_first.weight"], 1e-3)
    report("D conv2.weight_u after 3 forwards", d1["conv2.weight_u"], orc.d["conv2.weight_u"], 2e-3)
    check_fails()
