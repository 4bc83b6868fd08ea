"""Parity at the depth the benchmark runs at ("what's weak" #1).

  * test_full_step_23_blocks        one full optimize_parameters at 23 RRDBs (69 dense blocks, 702 G tensors), B = 4, against the
                                    fp32 CPU oracle (oracle/step.py = ssr_esrgan_model.py:119-233): every loss, the generator
                                    output, rel-L2 / cosine of EVERY gradient tensor of G and D against the plain oracle, the
                                    generator gradients again against the oracle evaluated with the engine's activation pattern
                                    (given the engine's own dL/d output), post-Adam update direction, EMA, spectral-norm u / v.
                                    Writes the per-tensor table to gpurun_out/parity/ (copied to profiles/ by hand).
  * test_trajectory_20_steps        20 consecutive steps, engine vs oracle loss curves; a second oracle run whose initial weights are
                                    perturbed by one fp32 ulp is the chaos / rounding noise band printed beside the engine's deviation.

Tolerances are MEASURED bounds with margin: bf16 operands, f32 accumulation.
"""
import json
import os
import statistics

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "parity")


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def cosine(got, ref):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    return (torch.dot(got, ref) / (got.norm() * ref.norm() + 1e-30)).item()


def nchw_mask(act_t, lo=0, hi=None):
    t = act_t[..., lo:hi] if hi is not None else act_t[..., lo:]
    return (t > 0).permute(0, 3, 1, 2).cpu()


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    lr = torch.randint(1, 256, (B, 24, 32, 32), generator=g, dtype=torch.uint8)
    hr = torch.randint(1, 256, (B, 3, 128, 128), generator=g, dtype=torch.uint8)
    hr = F.avg_pool2d(hr.float(), 5, 1, 2).round().clamp(0, 255).to(torch.uint8)   # USM mask not saturated everywhere
    return lr, hr


def _summary(vals):
    v = sorted(vals)
    return dict(max=v[-1], p95=v[int(0.95 * (len(v) - 1))], median=statistics.median(v), min=v[0], n=len(v))


def test_full_step_23_blocks():
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    B, nb = 4, 23
    gp = nets.rrdbnet_init(24, 3, num_block=nb, seed=21)
    dp = nets.unet_disc_init(27, seed=22)
    vp = losses.vgg19_init(seed=23)
    lr_u8, hr_u8 = _batch(B, 24)
    opt = dict(ema_decay=0.999, lr=1e-4)
    orc = OracleESRGAN(gp, dp, vp, opt, num_block=nb)
    orc.feed_data(lr_u8, hr_u8)
    ref_log = orc.optimize_parameters()

    tr = ESRGANTrainer(gp, dp, vp, dict(opt, network_g=dict(num_in_ch=24, num_block=nb)))
    tr.feed_data(lr_u8.pin_memory(), hr_u8.pin_memory())
    g0, d0 = tr.g_state_dict(), tr.d_state_dict()
    tr.optimize_parameters(1)
    log = tr.get_current_log()
    torch.cuda.synchronize()
    report = {"config": dict(num_block=nb, batch=B, dtype="bf16 operands / f32 accumulate"), "losses": {}}
    print()
    for k in ref_log:
        report["losses"][k] = dict(engine=log[k], oracle=ref_log[k])
        print(f"  {k:12s} engine {log[k]:+.6f} oracle {ref_log[k]:+.6f}")
        assert abs(log[k] - ref_log[k]) < 3e-2 * abs(ref_log[k]) + 2e-3, k
    e_out = rel_l2(tr.output, orc.output)
    report["output_rel_l2"] = e_out
    print(f"  generator output rel_l2 {e_out:.3e}")
    assert e_out < 1e-2

    # ---- every gradient tensor against the PLAIN oracle (kink flips of near-zero pre-activations included)
    rows = []
    for net, grads, ref in (("G", tr.g_grads(), orc.g), ("D", tr.d_grads(), orc.d)):
        for k, v in grads.items():
            rows.append(dict(net=net, name=k, numel=v.numel(), rel_l2_plain=rel_l2(v, ref[k].grad), cos_plain=cosine(v, ref[k].grad)))
    g_rows = [r for r in rows if r["net"] == "G"]
    d_rows = [r for r in rows if r["net"] == "D"]
    report["grad_plain"] = dict(G_rel_l2=_summary([r["rel_l2_plain"] for r in g_rows]), G_cos=_summary([r["cos_plain"] for r in g_rows]),
                                D_rel_l2=_summary([r["rel_l2_plain"] for r in d_rows]), D_cos=_summary([r["cos_plain"] for r in d_rows]))
    print("  plain oracle   G rel_l2", report["grad_plain"]["G_rel_l2"], "\n                 G cos   ", report["grad_plain"]["G_cos"])
    print("                 D rel_l2", report["grad_plain"]["D_rel_l2"], "\n                 D cos   ", report["grad_plain"]["D_cos"])

    # ---- generator backward against the oracle evaluated WITH the engine's LeakyReLU pattern, fed the engine's own dL/d output
    # (the network is piecewise linear: given the pattern and the output gradient, the parameter gradients are exact)
    ws = tr.G.workspace(B, 32, 32, True)
    masks = {}
    for i in range(nb):
        for j in range(3):
            buf = ws.bufs[3 * i + j].t
            for k in range(1, 5):
                lo = 64 + 32 * (k - 1)
                masks[f"body.{i}.rdb{j + 1}.conv{k}"] = nchw_mask(buf, lo, lo + 32)
    masks["conv_up1"], masks["conv_up2"], masks["conv_hr"] = nchw_mask(ws.up_out[0].t), nchw_mask(ws.up_out[1].t), nchw_mask(ws.hr.t)
    d_out = tr.io["d_out"].detach().float().cpu()          # dL/d output the engine back-propagated (L1 + perceptual + GAN)
    po = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    out_m = nets.rrdbnet_forward(po, orc.lr, num_block=nb, act=nets.masked_lrelu(masks))
    out_m.backward(d_out)
    by_name = {(r["net"], r["name"]): r for r in rows}
    forced = []
    for k, v in tr.g_grads().items():
        e, c = rel_l2(v, po[k].grad), cosine(v, po[k].grad)
        by_name[("G", k)].update(rel_l2_forced=e, cos_forced=c)
        forced.append(e)
    report["grad_pattern_forced_G"] = _summary(forced)
    print("  pattern-forced G rel_l2", report["grad_pattern_forced_G"])
    worst = max(forced)

    # ---- Adam update direction, EMA, spectral-norm state
    g1, d1 = tr.g_state_dict(), tr.d_state_dict()
    upd = {}
    for k in ("conv_first.weight", "body.0.rdb1.conv1.weight", "body.11.rdb2.conv3.weight", "body.22.rdb3.conv5.weight", "conv_last.weight"):
        upd[k] = cosine((g1[k] - g0[k]).cpu(), orc.g[k].detach() - gp[k])
    for k in ("conv0.weight", "conv3.weight_orig", "conv9.weight"):
        upd["D." + k] = cosine((d1[k] - d0[k]).cpu(), orc.d[k].detach() - dp[k])
    report["update_cosine"] = upd
    print("  update cosines", {k: round(v, 4) for k, v in upd.items()})
    ema_e = max(rel_l2(tr.g_state_dict(ema=True)[k], orc.g_ema[k]) for k in ("conv_first.weight", "body.22.rdb3.conv5.weight", "conv_last.weight"))
    uv_e = max(rel_l2(d1[k], orc.d[k]) for k in d1 if k.endswith(("weight_u", "weight_v")))
    report["ema_rel_l2"], report["sn_uv_rel_l2"] = ema_e, uv_e
    print(f"  EMA rel_l2 {ema_e:.2e}   spectral-norm u/v rel_l2 (after 3 forwards) {uv_e:.2e}")
    report["tensors"] = rows
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "full_step_nb23_b4.json"), "w") as fh:
        json.dump(report, fh, indent=1)

    # measured on the B200 (profiles/r02_parity_full_step_nb23_b4.json): pattern-forced max 7.5e-3; plain oracle G max 9.4e-3
    # (cosine >= 0.99996 on all 702 tensors), D max 2.6e-2; update cosines >= 0.986; EMA 3e-6; u / v 5e-7
    assert worst < 2e-2, f"pattern-forced generator gradients: worst rel_l2 {worst}"
    assert report["grad_plain"]["G_rel_l2"]["max"] < 3e-2 and report["grad_plain"]["G_cos"]["min"] > 0.999
    assert report["grad_plain"]["D_rel_l2"]["max"] < 6e-2 and report["grad_plain"]["D_cos"]["min"] > 0.999
    assert min(upd.values()) > 0.95
    assert ema_e < 1e-4 and uv_e < 1e-4


def test_trajectory_20_steps():
    """Loss curves over 20 steps at 23 blocks: |engine - oracle| per step and loss, next to the deviation of a second oracle run
    started one fp32 ulp away (how far a pure-fp32 implementation with another summation order would drift)."""
    from oracle import losses, nets
    from oracle.step import OracleESRGAN
    from satlas_super_resolution_b200.trainer import ESRGANTrainer
    B, nb, steps = 2, int(os.environ.get("SSR_TRAJ_BLOCKS", 23)), int(os.environ.get("SSR_TRAJ_STEPS", 20))
    gp = nets.rrdbnet_init(24, 3, num_block=nb, seed=31)
    dp = nets.unet_disc_init(27, seed=32)
    vp = losses.vgg19_init(seed=33)
    opt = dict(ema_decay=0.999, lr=1e-4)
    batches = [_batch(B, 100 + i) for i in range(steps)]
    gen = torch.Generator().manual_seed(5)
    ulp = lambda p: {k: (v * (1 + (torch.randint(0, 2, v.shape, generator=gen).float() * 2 - 1) * 2.0 ** -24)
                         if v.dtype.is_floating_point and not k.endswith(("weight_u", "weight_v")) else v.clone()) for k, v in p.items()}
    orc_a = OracleESRGAN(gp, dp, vp, opt, num_block=nb)
    orc_b = OracleESRGAN(ulp(gp), ulp(dp), vp, opt, num_block=nb)
    tr = ESRGANTrainer(gp, dp, vp, dict(opt, network_g=dict(num_in_ch=24, num_block=nb), cuda_graph=True))
    curves = []
    for i, (lr_u8, hr_u8) in enumerate(batches):
        orc_a.feed_data(lr_u8, hr_u8)
        la = orc_a.optimize_parameters(i + 1)
        orc_b.feed_data(lr_u8, hr_u8)
        lb = orc_b.optimize_parameters(i + 1)
        tr.feed_data(lr_u8, hr_u8)
        tr.optimize_parameters(i + 1)      # step 1 eager, step 2 captures the CUDA graph, the rest replay it
        le = tr.get_current_log()
        curves.append(dict(step=i + 1, oracle=la, oracle_ulp=lb, engine=dict(le)))
    torch.cuda.synchronize()
    print()
    keys = list(curves[0]["oracle"].keys())
    worst = {k: 0.0 for k in keys}
    band = {k: 0.0 for k in keys}
    for c in curves:
        row = []
        for k in keys:
            a, b, e = c["oracle"][k], c["oracle_ulp"][k], c["engine"][k]
            dev = abs(e - a) / (abs(a) + 1e-3)
            worst[k] = max(worst[k], dev)
            band[k] = max(band[k], abs(b - a) / (abs(a) + 1e-3))
            row.append(f"{k}={e:+.4f}/{a:+.4f}")
        print(f"  step {c['step']:2d}: " + "  ".join(row))
    print("  worst relative deviation engine vs oracle:", {k: f"{v:.2e}" for k, v in worst.items()})
    print("  fp32 one-ulp noise band (oracle vs oracle) :", {k: f"{v:.2e}" for k, v in band.items()})
    # weights after the trajectory: the update direction accumulated over 20 Adam steps
    g1 = tr.g_state_dict()
    cos = {k: cosine((g1[k].cpu() - gp[k]), orc_a.g[k].detach() - gp[k]) for k in ("conv_first.weight", "body.0.rdb1.conv1.weight", "conv_last.weight")}
    print("  cosine of the accumulated weight change (engine vs oracle):", {k: round(v, 4) for k, v in cos.items()})
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"trajectory_nb{nb}_b{B}.json"), "w") as fh:
        json.dump(dict(config=dict(num_block=nb, batch=B, steps=steps), curves=curves, worst_rel_dev=worst, ulp_band=band,
                       weight_change_cosine=cos), fh, indent=1)
    # Measured (profiles/r02_parity_trajectory_nb23_b2.json): the reconstruction losses stay within 1.2 % for all 20 steps and every loss
    # within 0.3 % for the first 10; after that the discriminator game amplifies ANY perturbation -- the two fp32 oracle runs that
    # start one ulp apart are 1.5 % apart in l_d_* and 100 % in out_d_* by step 20 (x1.8 per step) -- so the late GAN terms are
    # only bounded coarsely, next to that band.
    early = [c for c in curves if c["step"] <= 10]
    for k in keys:
        if k.startswith("out_d"):
            assert max(abs(c["engine"][k] - c["oracle"][k]) for c in early) < 0.01, k
            assert max(abs(c["engine"][k] - c["oracle"][k]) for c in curves) < 0.3, k
        else:
            assert max(abs(c["engine"][k] - c["oracle"][k]) / abs(c["oracle"][k]) for c in early) < 0.02, k
            assert worst[k] < (0.03 if k in ("l_g_pix", "l_g_percep") else 0.25), (k, worst[k])
    assert min(cos.values()) > 0.99
