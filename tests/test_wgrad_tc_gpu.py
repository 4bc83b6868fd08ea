"""GPU parity of the tcgen05 weight-gradient kernel (ssr_wgrad_tc + ssr_wgrad_unpack + ssr_bias_grad) vs torch CPU."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from test_conv_tc_gpu import bf16_round, nhwc_buffer, rel_err

pytestmark = pytest.mark.gpu


def run_wgrad(x, dy, r, scale=1.0, x_stride=None, dy_stride=None, dy_off=0, splits=0):
    from satlas_super_resolution_b200 import _lib as L
    from satlas_super_resolution_b200._protos import WgradArgs
    lib = L.load()
    B, cx, H, W = x.shape
    cy = dy.shape[1]
    xb = nhwc_buffer(x, x_stride)
    db = nhwc_buffer(dy, dy_stride, dy_off)
    cyp = (cy + 3) // 4 * 4
    acc = torch.zeros((r * r, cx, cyp), dtype=torch.float32, device="cuda")
    a = WgradArgs()
    a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cx = xb.data_ptr(), B, H, W, xb.shape[-1], cx
    a.dy, a.dy_pix_stride, a.cy, a.r = db.data_ptr() + 2 * dy_off, db.shape[-1], cy, r
    a.out, a.out_cx_rows, a.out_stride, a.scale, a.splits = acc.data_ptr(), cx, cyp, scale, splits
    L.check(lib.ssr_wgrad_tc(C.byref(a), None))
    grad = torch.full((cy, cx, r, r), 3.0, dtype=torch.float32, device="cuda")
    L.check(lib.ssr_wgrad_unpack(acc.data_ptr(), cx, cyp, grad.data_ptr(), cy, cx, r, 1.0, 0, None))
    bg = torch.zeros(cy, dtype=torch.float32, device="cuda")
    L.check(lib.ssr_bias_grad(db.data_ptr() + 2 * dy_off, db.shape[-1], B * H * W, cy, bg.data_ptr(), scale, None))
    torch.cuda.synchronize()
    return grad.cpu(), bg.cpu()


def ref_wgrad(x, dy, r, scale=1.0):
    w = torch.zeros(dy.shape[1], x.shape[1], r, r, requires_grad=True)
    b = torch.zeros(dy.shape[1], requires_grad=True)
    y = F.conv2d(x, w, b, padding=(r - 1) // 2)
    y.backward(dy * scale)
    return w.grad, b.grad


@pytest.mark.parametrize("B,cx,cy,H,W,r", [
    (1, 64, 32, 32, 32, 3),     # RDB conv1
    (2, 96, 32, 32, 32, 3),     # conv2: one 128-row M tile, rows 96.. ignored
    (2, 192, 64, 32, 32, 3),    # conv5: two M tiles, cy = 64 (128-byte dY rows)
    (1, 160, 32, 32, 32, 3),
    (1, 64, 64, 64, 64, 3),
    (1, 64, 3, 128, 128, 3),    # conv_last (dY padded to 16 channels)
    (2, 128, 128, 16, 16, 3),   # discriminator conv5-like: cy = 128 -> two dY blocks
    (1, 256, 256, 16, 16, 3),   # cy tiles in grid.z
    (1, 32, 64, 32, 32, 3),     # conv_first, channel-padded input
    (1, 1024, 128, 8, 64, 1),   # 1x1: the im2col GEMM form of a strided conv
])
def test_wgrad(B, cx, cy, H, W, r):
    g = torch.Generator().manual_seed(cx + cy + H)
    x = bf16_round(torch.randn(B, cx, H, W, generator=g))
    dy = bf16_round(torch.randn(B, cy, H, W, generator=g))
    got, gb = run_wgrad(x, dy, r, dy_stride=max(16, (cy + 7) // 8 * 8))
    ref, rb = ref_wgrad(x, dy, r)
    assert rel_err(got, ref) < 1e-4
    assert rel_err(gb, rb) < 1e-4


def test_wgrad_slices_and_scale():
    """X = channels [0,160) of a 192-wide dense-block buffer, dY = a 32-channel slice of a gradient buffer."""
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(2, 160, 32, 32, generator=g))
    dy = bf16_round(torch.randn(2, 32, 32, 32, generator=g))
    got, gb = run_wgrad(x, dy, 3, scale=0.2, x_stride=192, dy_stride=192, dy_off=160, splits=5)
    ref, rb = ref_wgrad(x, dy, 3, scale=0.2)
    assert rel_err(got, ref) < 1e-4
    assert rel_err(gb, rb) < 1e-4


@pytest.mark.parametrize("B,cx,cy,H,W", [(2, 64, 128, 32, 32), (1, 128, 256, 64, 64), (2, 256, 512, 16, 32), (4, 64, 128, 128, 128)])
def test_wgrad_strided_4x4(B, cx, cy, H, W):
    """weight gradient of the 4 x 4 stride-2 pad-1 conv straight from the NHWC input (element-strided TMA gather per tap)"""
    from satlas_super_resolution_b200 import _lib as L
    from satlas_super_resolution_b200._protos import WgradArgs
    lib = L.load()
    g = torch.Generator().manual_seed(cx + H)
    x = bf16_round(torch.randn(B, cx, H, W, generator=g))
    dy = bf16_round(torch.randn(B, cy, H // 2, W // 2, generator=g))
    xb, db = nhwc_buffer(x), nhwc_buffer(dy)
    acc = torch.zeros((16, cx, cy), dtype=torch.float32, device="cuda")
    a = WgradArgs()
    a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cx = xb.data_ptr(), B, H, W, cx, cx
    a.dy, a.dy_pix_stride, a.cy, a.r = db.data_ptr(), cy, cy, 4
    a.out, a.out_cx_rows, a.out_stride, a.scale, a.splits = acc.data_ptr(), cx, cy, 1.0, 0
    L.check(lib.ssr_wgrad_tc(C.byref(a), None))
    grad = torch.zeros((cy, cx, 4, 4), dtype=torch.float32, device="cuda")
    L.check(lib.ssr_wgrad_unpack(acc.data_ptr(), cx, cy, grad.data_ptr(), cy, cx, 4, 1.0, 0, None))
    torch.cuda.synchronize()
    w = torch.zeros(cy, cx, 4, 4, requires_grad=True)
    F.conv2d(x, w, stride=2, padding=1).backward(dy)
    assert rel_err(grad.cpu(), w.grad) < 1e-4
