"""Restatement of the basicsr==1.4.2 pieces the step uses (basicsr is a requirements.txt:1 dependency that is not
vendored under /root/reference and not installed here -- "parity unpinned" by the reference; every function below is
cross-checked in tests/test_oracle.py against the torch / torchvision / cv2 primitive it wraps).

Reference call sites: /root/reference/ssr/models/ssr_esrgan_model.py:31,109 (USMSharp), :148 (L1Loss),
:154 (PerceptualLoss), :182,218,224 (GANLoss); configuration ssr/options/esrgan_s2naip_urban.yml:118-144.
"""
import math

import torch
import torch.nn.functional as F

VGG19_LAYERS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool1", ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool2",
                ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), "pool3",
                ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512), "pool4",
                ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512)]
DEFAULT_LAYER_WEIGHTS = {"conv1_2": 0.1, "conv2_2": 0.1, "conv3_4": 1.0, "conv4_4": 1.0, "conv5_4": 1.0}
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def l1_loss(pred, target, loss_weight=1.0):
    """basicsr L1Loss(loss_weight, reduction='mean')."""
    return loss_weight * F.l1_loss(pred, target, reduction="mean")


def gan_loss_vanilla(pred, target_is_real, is_disc, loss_weight=1.0, real_label_val=1.0, fake_label_val=0.0):
    """basicsr GANLoss('vanilla'): BCEWithLogits against a constant label map; the weight applies to the generator only."""
    label = real_label_val if target_is_real else fake_label_val
    loss = F.binary_cross_entropy_with_logits(pred, torch.full_like(pred, label))
    return loss if is_disc else loss * loss_weight


def vgg19_init(seed=0):
    """Seeded-random VGG19 feature weights (the pretrained file is unobtainable offline) with
    torchvision's own init: kaiming_normal(fan_out, relu), zero bias."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for c in VGG19_LAYERS:
        if isinstance(c, str):
            continue
        name, cin, cout = c
        std = math.sqrt(2.0 / (cout * 9))
        p[f"{name}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
        p[f"{name}.bias"] = torch.zeros(cout)
    return p


def masked_relu_pool(relu_masks, pool_index):
    """(relu_fn, pool_fn) whose active sets / arg-max choices are DICTATED by the engine's saved activations (see
    oracle.nets.masked_lrelu): relu_masks[name] bool [N,C,H,W]; pool_index[name] int64 [N,C,H/2,W/2] in 0..3."""
    def relu_fn(name, t):
        return t * relu_masks[name].to(t.dtype)

    def pool_fn(name, t):
        n, c, h, w = t.shape
        win = t.view(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
        return torch.gather(win, 4, pool_index[name].unsqueeze(-1)).squeeze(-1)
    return relu_fn, pool_fn


def vgg19_features(p, x, layer_names, use_input_norm=True, range_norm=False, relu_fn=None, pool_fn=None):
    """basicsr VGGFeatureExtractor(vgg19): pre-ReLU conv outputs by name; MaxPool2d(2,2) kept."""
    relu_fn = relu_fn or (lambda name, t: F.relu(t))
    pool_fn = pool_fn or (lambda name, t: F.max_pool2d(t, 2, 2))
    if range_norm:
        x = (x + 1) / 2
    if use_input_norm:
        mean = torch.tensor(IMAGENET_MEAN, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, device=x.device).view(1, 3, 1, 1)
        x = (x - mean) / std
    out = {}
    remaining = set(layer_names)
    for c in VGG19_LAYERS:
        if isinstance(c, str):
            x = pool_fn(c, x)
            continue
        name = c[0]
        x = F.conv2d(x, p[f"{name}.weight"], p[f"{name}.bias"], padding=1)
        if name in remaining:
            out[name] = x
            remaining.discard(name)
            if not remaining:
                break
        x = relu_fn(name, x)
    return out


def perceptual_loss(p, x, gt, layer_weights=None, perceptual_weight=1.0, use_input_norm=True, range_norm=False,
                    relu_fn=None, pool_fn=None):
    """basicsr PerceptualLoss(criterion='l1', style_weight=0): sum_k w_k * L1(vgg_k(x), vgg_k(gt.detach()))."""
    lw = layer_weights or DEFAULT_LAYER_WEIGHTS
    fx = vgg19_features(p, x, lw.keys(), use_input_norm, range_norm, relu_fn, pool_fn)
    with torch.no_grad():
        fg = vgg19_features(p, gt.detach(), lw.keys(), use_input_norm, range_norm)
    loss = 0
    for k, w in lw.items():
        loss = loss + F.l1_loss(fx[k], fg[k]) * w
    return loss * perceptual_weight


def gaussian_kernel_1d(ksize=51, sigma=0.0):
    """cv2.getGaussianKernel(ksize, sigma): sigma <= 0 -> 0.3*((ksize-1)*0.5 - 1) + 0.8 (= 8.0 for 51)."""
    if sigma <= 0:
        sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = torch.arange(ksize, dtype=torch.float64) - (ksize - 1) / 2
    k = torch.exp(-(x * x) / (2 * sigma * sigma))
    return (k / k.sum())


def filter2d(img, kernel2d):
    """basicsr filter2D: reflect pad, depthwise cross-correlation with a k x k kernel."""
    k = kernel2d.shape[-1]
    b, c, h, w = img.shape
    x = F.pad(img, (k // 2, k // 2, k // 2, k // 2), mode="reflect")
    x = x.view(b * c, 1, x.shape[-2], x.shape[-1])
    return F.conv2d(x, kernel2d.view(1, 1, k, k)).view(b, c, h, w)


def usm_sharp(img, radius=50, sigma=0, weight=0.5, threshold=10):
    """basicsr USMSharp(radius=50, sigma=0).forward(img, weight=0.5, threshold=10)."""
    if radius % 2 == 0:
        radius += 1
    k1 = gaussian_kernel_1d(radius, sigma)
    kernel = torch.outer(k1, k1).to(torch.float32).to(img.device)
    blur = filter2d(img, kernel)
    residual = img - blur
    mask = (residual.abs() * 255 > threshold).float()
    soft_mask = filter2d(mask, kernel)
    sharp = torch.clip(img + weight * residual, 0, 1)
    return soft_mask * sharp + (1 - soft_mask) * img


def ssim_loss(x, gt, loss_weight=1.0, window_size=5, sigma=1.5, max_val=1.0, eps=1e-12):
    """SSIMLoss -- /root/reference/ssr/losses/basic_loss.py:50-60 (call site ssr_esrgan_model.py:163-164):
    `kornia.losses.ssim_loss(x, gt, window_size=5, reduction="none")`, mean over (C, H, W), mean over the batch, times loss_weight.
    kornia is a requirements.txt:7 dependency (no version pinned) that is not installed here -- "parity unpinned"; restated from
    kornia.metrics.ssim / kornia.losses.ssim_loss: Gaussian window get_gaussian_kernel1d(5, 1.5), filter2d with reflect padding
    ('same'), C1 = (0.01 max_val)^2, C2 = (0.03 max_val)^2, ssim = num / (den + eps), loss map = clamp((1 - ssim) / 2, 0, 1)."""
    k1 = gaussian_kernel_1d(window_size, sigma).to(torch.float32)
    kernel = torch.outer(k1, k1).to(x.device)
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu1, mu2 = filter2d(x, kernel), filter2d(gt, kernel)
    s11 = filter2d(x * x, kernel) - mu1 * mu1
    s22 = filter2d(gt * gt, kernel) - mu2 * mu2
    s12 = filter2d(x * gt, kernel) - mu1 * mu2
    num = (2.0 * mu1 * mu2 + c1) * (2.0 * s12 + c2)
    den = (mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2)
    loss_map = torch.clamp((1.0 - num / (den + eps)) / 2, min=0, max=1)
    return torch.mean(loss_map.mean(dim=(-1, -2, -3))) * loss_weight
