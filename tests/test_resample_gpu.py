"""GPU parity of the resampling kernels (F.interpolate nearest / bilinear x2 and their gradients) against torch CPU fp32.

The kernels work on bf16 NHWC buffers; the oracle is F.interpolate on the SAME bf16-rounded values in fp32, so the only
difference left is the final bf16 rounding: tolerance 2^-8 of the output scale (reference: discriminator_arch.py:50-60,
rrdbnet_arch.py:124-131)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from satlas_super_resolution_b200 import _lib as L
    return L, L.load()


def nhwc(t, stride=None):
    """[B,C,H,W] cpu f32 -> bf16 NHWC cuda buffer with the given pixel stride"""
    B, Cc, H, W = t.shape
    buf = torch.full((B, H, W, stride or Cc), 3.0, dtype=torch.bfloat16, device="cuda")
    buf[..., :Cc] = t.permute(0, 2, 3, 1).to("cuda", torch.bfloat16)
    return buf


def back(buf, Cc):
    return buf[..., :Cc].float().cpu().permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(torch.bfloat16).float()


@pytest.mark.parametrize("B,Cc,H,W,skip", [(2, 64, 16, 16, True), (3, 128, 8, 24, False), (1, 8, 1, 5, True), (2, 512, 4, 4, True),
                                           (1, 16, 33, 7, False)])
def test_bilinear2x_forward_matches_interpolate(B, Cc, H, W, skip):
    L, lib = _lib()
    x = rnd(B, Cc, H, W, seed=H + W)
    s = rnd(B, Cc, H, W, seed=H * W + 1) if skip else None
    xb, sb = nhwc(x, Cc + 8), (nhwc(s) if skip else None)
    ob = torch.zeros((B, 2 * H, 2 * W, Cc + 16), dtype=torch.bfloat16, device="cuda")
    L.check(lib.ssr_upsample_bilinear2x(xb.data_ptr(), xb.shape[-1], sb.data_ptr() if skip else None, sb.shape[-1] if skip else 0,
                                        ob.data_ptr(), ob.shape[-1], B, H, W, Cc, None))
    torch.cuda.synchronize()
    src = x + s if skip else x      # the skip is added in f32 inside the kernel; only the interpolated result is rounded to bf16
    ref = F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=False)
    got = back(ob, Cc)
    assert (got - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    assert torch.all(ob[..., Cc:] == 0)          # pad channels of the output buffer untouched


@pytest.mark.parametrize("B,Cc,H,W", [(2, 64, 16, 16), (3, 128, 8, 24), (1, 8, 1, 5), (2, 256, 4, 4), (1, 16, 33, 7)])
def test_bilinear2x_backward_matches_autograd(B, Cc, H, W):
    L, lib = _lib()
    dy = rnd(B, Cc, 2 * H, 2 * W, seed=H + 3 * W)
    x = torch.zeros(B, Cc, H, W, requires_grad=True)
    F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False).backward(dy)
    dyb = nhwc(dy, Cc + 8)
    dxb = torch.zeros((B, H, W, Cc), dtype=torch.bfloat16, device="cuda")
    L.check(lib.ssr_upsample_bilinear2x_bwd(dyb.data_ptr(), dyb.shape[-1], dxb.data_ptr(), dxb.shape[-1], B, H, W, Cc, None))
    torch.cuda.synchronize()
    got = back(dxb, Cc)
    assert (got - x.grad).abs().max().item() <= 2 ** -7 * x.grad.abs().max().item()
