"""Functional fp32 restatement of the two reference networks over a plain {name: tensor} state dict.

Key names are exactly the reference state_dict schema:
  G: conv_first, body.<i>.rdb<1-3>.conv<1-5>, conv_body, conv_up1, conv_up2, conv_hr, conv_last (.weight/.bias)
  D: conv0.{weight,bias}, conv<1-8>.{weight_orig,weight_u,weight_v}, conv9.{weight,bias}
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- generator
def _lrelu(name, t):
    return F.leaky_relu(t, 0.2)


def masked_lrelu(masks):
    """LeakyReLU(0.2) whose active set is DICTATED by `masks[name]` (bool, same shape) instead of sign(t): the network
    becomes the piecewise-linear map with the engine's activation pattern, so its autograd gradients are exactly what
    the engine's backward kernels must produce (used by tests/test_train_gpu.py; values differ from the true
    LeakyReLU only where |t| is within rounding noise of zero)."""
    def act(name, t):
        m = masks[name].to(t.dtype)
        return t * (m + (1 - m) * 0.2)
    return act


def rdb_forward(p, pre, x, act=_lrelu):
    """ResidualDenseBlock.forward -- /root/reference/ssr/archs/rrdbnet_arch.py:37-44."""
    feats = [x]
    for k in range(1, 5):
        y = F.conv2d(torch.cat(feats, 1), p[f"{pre}.conv{k}.weight"], p[f"{pre}.conv{k}.bias"], padding=1)
        feats.append(act(f"{pre}.conv{k}", y))
    x5 = F.conv2d(torch.cat(feats, 1), p[f"{pre}.conv5.weight"], p[f"{pre}.conv5.bias"], padding=1)
    return x5 * 0.2 + x


def rrdb_forward(p, pre, x, act=_lrelu):
    """RRDB.forward -- rrdbnet_arch.py:63-68."""
    out = x
    for j in (1, 2, 3):
        out = rdb_forward(p, f"{pre}.rdb{j}", out, act)
    return out * 0.2 + x


def pixel_unshuffle(x, scale):
    """arch_util.py:769-785 (only reached for scale 1 / 2)."""
    b, c, hh, hw = x.shape
    h, w = hh // scale, hw // scale
    return x.view(b, c, h, scale, w, scale).permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


def rrdbnet_forward(p, x, scale=4, num_block=23, act=_lrelu):
    """SSR_RRDBNet.forward -- rrdbnet_arch.py:116-137."""
    if scale == 2:
        feat = pixel_unshuffle(x, 2)
    elif scale == 1:
        feat = pixel_unshuffle(x, 4)
    else:
        feat = x
    conv = lambda name, t: F.conv2d(t, p[f"{name}.weight"], p[f"{name}.bias"], padding=1)
    feat = conv("conv_first", feat)
    body = feat
    for i in range(num_block):
        body = rrdb_forward(p, f"body.{i}", body, act)
    feat = feat + conv("conv_body", body)
    feat = act("conv_up1", conv("conv_up1", F.interpolate(feat, scale_factor=2, mode="nearest")))
    feat = act("conv_up2", conv("conv_up2", F.interpolate(feat, scale_factor=2, mode="nearest")))
    if scale in (8, 16):
        feat = act("conv_up3", conv("conv_up3", F.interpolate(feat, scale_factor=2, mode="nearest")))
        if scale == 16:
            feat = act("conv_up4", conv("conv_up4", F.interpolate(feat, scale_factor=2, mode="nearest")))
    return conv("conv_last", act("conv_hr", conv("conv_hr", feat)))


def rrdbnet_init(num_in_ch, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=4, seed=0):
    """Weights with the reference's init distributions (rrdbnet_arch.py:35,99-112; arch_util.py:600-628):
    RDB convs kaiming_normal * 0.1 with zero bias, the rest nn.Conv2d's default init.  Deterministic in `seed`
    (own generator, independent of the order nn.Module construction would consume the global RNG in)."""
    g = torch.Generator().manual_seed(seed)
    if scale == 2:
        num_in_ch *= 4
    elif scale == 1:
        num_in_ch *= 16
    p = {}

    def default_conv(name, cout, cin, k=3):
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        p[f"{name}.weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        p[f"{name}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def rdb_conv(name, cout, cin):
        std = math.sqrt(2.0 / (cin * 9))
        p[f"{name}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std * 0.1
        p[f"{name}.bias"] = torch.zeros(cout)

    default_conv("conv_first", num_feat, num_in_ch)
    for i in range(num_block):
        for j in (1, 2, 3):
            for k in range(1, 5):
                rdb_conv(f"body.{i}.rdb{j}.conv{k}", num_grow_ch, num_feat + (k - 1) * num_grow_ch)
            rdb_conv(f"body.{i}.rdb{j}.conv5", num_feat, num_feat + 4 * num_grow_ch)
    default_conv("conv_body", num_feat, num_feat)
    default_conv("conv_up1", num_feat, num_feat)
    default_conv("conv_up2", num_feat, num_feat)
    if scale in (8, 16):
        default_conv("conv_up3", num_feat, num_feat)
        if scale == 16:
            default_conv("conv_up4", num_feat, num_feat)
    default_conv("conv_hr", num_feat, num_feat)
    default_conv("conv_last", num_out_ch, num_feat)
    return p


# --------------------------------------------------------------------------- discriminator
def _normalize(v, eps=1e-12):
    return v / v.norm().clamp_min(eps)


def spectral_norm_weight(p, name, training, update_state=True):
    """torch.nn.utils.spectral_norm (legacy hook), n_power_iterations=1, dim=0, eps=1e-12.

    Training mode: ONE in-place power iteration on (u, v) under no_grad on every forward (also when the
    weights are frozen), then sigma = u^T W v with u, v constants; weight = weight_orig / sigma.
    Eval mode: no iteration.  Used at /root/reference/ssr/archs/discriminator_arch.py:30-39."""
    w = p[f"{name}.weight_orig"]
    wm = w.reshape(w.shape[0], -1)
    u, v = p[f"{name}.weight_u"], p[f"{name}.weight_v"]
    if training:
        with torch.no_grad():
            v_new = _normalize(torch.mv(wm.t(), u))
            u_new = _normalize(torch.mv(wm, v_new))
        if update_state:
            p[f"{name}.weight_u"] = u_new
            p[f"{name}.weight_v"] = v_new
        u, v = u_new, v_new
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def unet_disc_forward(p, x, training=True, skip_connection=True, update_state=True, act=_lrelu):
    """SSR_UNetDiscriminatorSN.forward -- discriminator_arch.py:42-71."""
    sn = lambda name: spectral_norm_weight(p, name, training, update_state)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    x0 = act("conv0", F.conv2d(x, p["conv0.weight"], p["conv0.bias"], padding=1))
    x1 = act("conv1", F.conv2d(x0, sn("conv1"), None, stride=2, padding=1))
    x2 = act("conv2", F.conv2d(x1, sn("conv2"), None, stride=2, padding=1))
    x3 = act("conv3", F.conv2d(x2, sn("conv3"), None, stride=2, padding=1))
    x3 = up(x3)
    x4 = act("conv4", F.conv2d(x3, sn("conv4"), None, padding=1))
    if skip_connection:
        x4 = x4 + x2
    x4 = up(x4)
    x5 = act("conv5", F.conv2d(x4, sn("conv5"), None, padding=1))
    if skip_connection:
        x5 = x5 + x1
    x5 = up(x5)
    x6 = act("conv6", F.conv2d(x5, sn("conv6"), None, padding=1))
    if skip_connection:
        x6 = x6 + x0
    out = act("conv7", F.conv2d(x6, sn("conv7"), None, padding=1))
    out = act("conv8", F.conv2d(out, sn("conv8"), None, padding=1))
    return F.conv2d(out, p["conv9.weight"], p["conv9.bias"], padding=1)


def unet_disc_init(num_in_ch, num_feat=64, seed=0):
    """Default nn.Conv2d init + spectral_norm's u/v init (normal, normalised), deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    nf = num_feat
    specs = [("conv0", nf, num_in_ch, 3, True), ("conv1", nf * 2, nf, 4, False), ("conv2", nf * 4, nf * 2, 4, False),
             ("conv3", nf * 8, nf * 4, 4, False), ("conv4", nf * 4, nf * 8, 3, False), ("conv5", nf * 2, nf * 4, 3, False),
             ("conv6", nf, nf * 2, 3, False), ("conv7", nf, nf, 3, False), ("conv8", nf, nf, 3, False),
             ("conv9", 1, nf, 3, True)]
    for name, cout, cin, k, has_bias in specs:
        bound = 1.0 / math.sqrt(cin * k * k)
        w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        if has_bias:
            p[f"{name}.weight"] = w
            p[f"{name}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        else:
            p[f"{name}.weight_orig"] = w
            p[f"{name}.weight_u"] = _normalize(torch.randn(cout, generator=g))
            p[f"{name}.weight_v"] = _normalize(torch.randn(cin * k * k, generator=g))
    return p


G_PARAM_COUNT_RGB8 = 16_710_083   # SURVEY.md [probe]: num_in_ch=24
D_PARAM_COUNT_RGB8 = 4_390_721    # num_in_ch=27


# --------------------------------------------------------------------------- rounding model of the engine's forward
def rrdbnet_forward_bf16_model(p, x, scale=4, num_block=23):
    """The generator forward with bf16 rounding applied at exactly the points the B200 engine rounds:
    conv OPERANDS (stored activations and weights) are bf16, every accumulation, bias, residual add and the 64-channel trunk are
    f32.  Evaluated in fp32 on the CPU, so what is left between this model and the engine is f32 summation order plus the
    occasional bf16 neighbour flip it causes -- a kernel bug is NOT absorbed by this model, bf16 operand noise is.
    (Test infrastructure: separates "kernel wrong" from "bf16 noise" without a second operand precision in the kernels.)"""
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    w = {k: (bf(v) if k.endswith(".weight") else v) for k, v in p.items()}
    conv = lambda name, t: F.conv2d(t, w[f"{name}.weight"], w[f"{name}.bias"], padding=1)     # t is already bf16-valued
    if scale == 2:
        x = pixel_unshuffle(x, 2)
    elif scale == 1:
        x = pixel_unshuffle(x, 4)
    trunk = conv("conv_first", bf(x))          # f32 trunk; its bf16 copy is the operand of the first dense block
    trunk0 = trunk
    for i in range(num_block):
        rrdb_in = trunk
        for j in (1, 2, 3):
            pre = f"body.{i}.rdb{j}"
            feats = [bf(trunk)]
            for k in range(1, 5):
                feats.append(bf(F.leaky_relu(conv(f"{pre}.conv{k}", torch.cat(feats, 1)), 0.2)))
            x5 = conv(f"{pre}.conv5", torch.cat(feats, 1))
            if j < 3:
                trunk = x5 * 0.2 + trunk                          # epilogue: s0 * (acc + bias) + s1 * trunk, all f32
            else:
                trunk = x5 * 0.04 + trunk * 0.2 + rrdb_in         # (x5 * 0.2 + x_rdb3) * 0.2 + x_rrdb folded into one epilogue
    feat = bf(conv("conv_body", bf(trunk)) + trunk0)
    n_up = {1: 2, 2: 2, 4: 2, 8: 3, 16: 4}[scale]
    for u in range(1, n_up + 1):
        feat = bf(F.leaky_relu(conv(f"conv_up{u}", F.interpolate(feat, scale_factor=2, mode="nearest")), 0.2))
    feat = bf(F.leaky_relu(conv("conv_hr", feat), 0.2))
    return conv("conv_last", feat)
