"""Registry-facing loss modules with the call shapes the reference step uses
(/root/reference/ssr/models/ssr_esrgan_model.py:148,154,182,218,224; basicsr L1Loss / GANLoss / PerceptualLoss):

    cri_pix(pred, target) -> scalar
    cri_perceptual(x, gt) -> (percep | None, style | None)
    cri_gan(pred, target_is_real: bool, is_disc: bool = False) -> scalar
    ssim_loss(x, gt) -> scalar                      (ssr/losses/basic_loss.py:50-60, call site ssr_esrgan_model.py:163-164)

Each is an autograd Function over kernels of libssr_b200 (loss value + gradient in one pass); CUDA tensors only.
"""
import torch
from torch import nn

from . import _lib as L
from . import weights
from .ops import cur_stream, lib
from .registry import LOSS_REGISTRY, _register


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weight):
        need = ctx.needs_input_grad[0]          # of the caller's tensor: the casts below return plain copies inside forward
        ctx.in_dtype = pred.dtype
        pred, target = pred.contiguous().float(), target.contiguous().float()
        loss = torch.zeros(1, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred) if need else None
        L.check(lib().ssr_l1_loss(pred.data_ptr(), target.data_ptr(), pred.numel(), weight, loss.data_ptr(),
                                  grad.data_ptr() if grad is not None else None, 0, cur_stream()))
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return ((ctx.grad * g).to(ctx.in_dtype) if ctx.grad is not None else None), None, None


class L1Loss(nn.Module):
    """basicsr L1Loss(loss_weight=1.0, reduction='mean')"""

    def __init__(self, loss_weight=1.0, reduction="mean"):
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("L1Loss: only reduction='mean' (what every shipped config uses) is built")
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, **kwargs):
        if weight is not None:
            raise NotImplementedError("L1Loss: element-wise weights are not built")
        return _L1Fn.apply(pred, target.detach(), float(self.loss_weight))


class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gt, weight):
        need = ctx.needs_input_grad[0]
        ctx.in_dtype = x.dtype
        x, gt = x.contiguous().float(), gt.contiguous().float()
        if x.dim() != 4 or x.shape != gt.shape:
            raise ValueError(f"SSIMLoss: expected two [B, C, H, W] tensors of one shape, got {tuple(x.shape)} / {tuple(gt.shape)}")
        B, C, H, W = x.shape
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if need else None
        scratch = torch.empty((3,) + tuple(x.shape), dtype=torch.float32, device=x.device) if need else None
        L.check(lib().ssr_ssim_loss(x.data_ptr(), gt.data_ptr(), B * C, H, W, weight, loss.data_ptr(),
                                    grad.data_ptr() if need else None, 0, scratch.data_ptr() if need else None, cur_stream()))
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return ((ctx.grad * g).to(ctx.in_dtype) if ctx.grad is not None else None), None, None


class SSIMLoss(nn.Module):
    """ssr/losses/basic_loss.py:50-60: kornia.losses.ssim_loss(x, gt, window_size=5, reduction='none'), mean over (C, H, W), mean
    over the batch, times loss_weight -- one forward and one adjoint-filter kernel of libssr_b200 (ssr_ssim_loss)."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, x, gt):
        return _SsimFn.apply(x, gt.detach(), float(self.loss_weight))


class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, weight):
        x = x.contiguous().float()
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        L.check(lib().ssr_bce_logits(x.data_ptr(), x.numel(), target, weight, loss.data_ptr(), None, grad.data_ptr(), cur_stream()))
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return ctx.grad * g, None, None


class GANLoss(nn.Module):
    """basicsr GANLoss: 'vanilla' = BCEWithLogitsLoss against a constant label map; loss_weight applies when is_disc=False."""

    def __init__(self, gan_type, real_label_val=1.0, fake_label_val=0.0, loss_weight=1.0):
        super().__init__()
        if gan_type != "vanilla":
            raise NotImplementedError(f"GANLoss: gan_type '{gan_type}' is not built (every shipped config uses 'vanilla')")
        self.gan_type, self.loss_weight = gan_type, loss_weight
        self.real_label_val, self.fake_label_val = real_label_val, fake_label_val

    def forward(self, input, target_is_real, is_disc=False):
        label = self.real_label_val if target_is_real else self.fake_label_val
        return _BceFn.apply(input, float(label), 1.0 if is_disc else float(self.loss_weight))


class _PercepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gt, mod):
        need = ctx.needs_input_grad[0]
        ctx.in_dtype = x.dtype
        x, gt = x.contiguous().float(), gt.contiguous().float()
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        dx = torch.zeros_like(x) if need else None
        mod.engine(x.device).loss_and_grad(x, gt, loss, dx, cur_stream())
        ctx.dx = dx
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return ((ctx.dx * g).to(ctx.in_dtype) if ctx.dx is not None else None), None, None


class PerceptualLoss(nn.Module):
    """basicsr PerceptualLoss(layer_weights, vgg_type='vgg19', use_input_norm, range_norm, perceptual_weight, style_weight=0,
    criterion='l1').  VGG19 weights: the torchvision checkpoint vgg19-dcbb9e9d.pth (weights.vgg19_search_paths()); when it is
    absent construction of the engine RAISES unless `vgg_seed` (or $SSR_VGG_RANDOM_SEED) explicitly asks for seeded random
    weights (tests / benchmarks offline)."""

    def __init__(self, layer_weights, vgg_type="vgg19", use_input_norm=True, range_norm=False, perceptual_weight=1.0,
                 style_weight=0.0, criterion="l1", vgg_seed=None, vgg_path=None):
        super().__init__()
        if vgg_type != "vgg19" or criterion != "l1" or style_weight:
            raise NotImplementedError("PerceptualLoss: only vgg19 / l1 / style_weight=0 (the shipped config) is built")
        self.layer_weights, self.perceptual_weight = dict(layer_weights), perceptual_weight
        self.use_input_norm, self.range_norm, self.vgg_seed, self.vgg_path = use_input_norm, range_norm, vgg_seed, vgg_path
        self._engine = None

    def engine(self, device):
        if self._engine is None:
            from .vgg import PerceptualEngine
            sd = weights.resolve_vgg19_state(self.vgg_seed, self.vgg_path)
            self._engine = PerceptualEngine({k: v.to(device) for k, v in sd.items()}, self.layer_weights, self.perceptual_weight,
                                            self.use_input_norm, self.range_norm)
        return self._engine

    def forward(self, x, gt):
        if self.perceptual_weight <= 0:
            return None, None
        return _PercepFn.apply(x, gt.detach(), self), None


for _c in (L1Loss, GANLoss, PerceptualLoss, SSIMLoss):
    _register(LOSS_REGISTRY, _c)
