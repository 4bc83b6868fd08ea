"""GPU parity of the tcgen05 implicit-GEMM conv (ssr_conv_tc through the C ABI) against torch CPU fp32.

Oracle = F.conv2d in fp32 on the CPU fed the SAME bf16-rounded activations and weights, so the only
differences left are fp32 accumulation order and the final bf16 rounding of the output:
tolerance = 2^-8 relative to the output scale for bf16 outputs, 1e-4 for f32 outputs.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from satlas_super_resolution_b200 import _lib as L
    return L, L.load()


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc_buffer(x_nchw, pix_stride=None, ch_off=0, fill=0.0):
    """[B,C,H,W] f32 cpu -> cuda bf16 buffer [B,H,W,pix_stride] holding x at channels [ch_off, ch_off+C)."""
    B, Cc, H, W = x_nchw.shape
    ps = pix_stride or Cc
    buf = torch.full((B, H, W, ps), fill, dtype=torch.bfloat16, device="cuda")
    buf[..., ch_off:ch_off + Cc] = x_nchw.permute(0, 2, 3, 1).to("cuda", torch.bfloat16)
    return buf


def pack_weight(L, lib, w, mode, k_pad=None):
    cout, cin, r, _ = w.shape
    red, outc = (cin, cout) if mode == L.PACK_FWD else (cout, cin)
    n_pad = C.c_int32(0)
    kp = k_pad or ((red + 63) // 64 * 64)
    nbytes = lib.ssr_packed_weight_bytes(kp, outc, r, C.byref(n_pad))
    packed = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    wd = w.to("cuda", torch.float32).contiguous()
    L.check(lib.ssr_pack_conv_weight(wd.data_ptr(), cout, cin, r, mode, None, packed.data_ptr(), kp, n_pad.value, None))
    return packed, n_pad.value


def run_conv(x, w, bias=None, act=0, s0=1.0, res1=None, s1=0.0, res1_f32=False, res2=None, s2=0.0,
             mask=None, mask_lo=0, mask_relu=0, out_kind="bf16", in_stride=None, in_off=0, cin_read=None,
             out_stride=None, out_off=0, mt=0, n_tile=0, splits=0, mode=None):
    """x [B,Cin,H,W], w [Cout,Cin,R,R] (cpu f32) -> output [B,Cout',H,W] cpu f32 computed by the library."""
    L, lib = _lib()
    mode = L.PACK_FWD if mode is None else mode
    B, Cin, H, W = x.shape
    r = w.shape[-1]
    outc = w.shape[0] if mode == L.PACK_FWD else w.shape[1]
    xb = nhwc_buffer(x, in_stride, in_off)
    ps = xb.shape[-1]
    cin = cin_read or Cin
    cin16 = (cin + 15) // 16 * 16
    packed, n_pad = pack_weight(L, lib, w, mode)
    a = L.ConvTcArgs()
    a.x = xb.data_ptr() + in_off * 2
    a.n_img, a.h, a.w = B, H, W
    a.x_pix_stride = ps
    a.cin = cin16
    a.w_packed = packed.data_ptr()
    a.r, a.cout, a.n_pad = r, outc, n_pad
    keep = [xb, packed]
    if bias is not None:
        bd = bias.to("cuda", torch.float32).contiguous()
        keep.append(bd)
        a.bias = bd.data_ptr()
    a.act = act
    a.s0 = s0
    if res1 is not None:
        if res1_f32:
            r1 = res1.permute(0, 2, 3, 1).contiguous().to("cuda", torch.float32)
            a.res1_kind = L.SSR_F32
        else:
            r1 = nhwc_buffer(res1)
            a.res1_kind = L.SSR_BF16
        keep.append(r1)
        a.res1 = r1.data_ptr()
        a.res1_pix_stride = r1.shape[-1]
        a.s1 = s1
    if res2 is not None:
        r2 = nhwc_buffer(res2)
        keep.append(r2)
        a.res2 = r2.data_ptr()
        a.res2_kind = L.SSR_BF16
        a.res2_pix_stride = r2.shape[-1]
        a.s2 = s2
    if mask is not None:
        mk = nhwc_buffer(mask)
        keep.append(mk)
        a.mask = mk.data_ptr()
        a.mask_pix_stride = mk.shape[-1]
        a.mask_lo = mask_lo
        a.mask_relu = mask_relu
    ostride = out_stride or outc
    if out_kind == "bf16":
        ob = torch.full((B, H, W, ostride), 7.0, dtype=torch.bfloat16, device="cuda")
        a.out_bf16 = ob.data_ptr() + out_off * 2
        a.out_pix_stride = ostride
    elif out_kind == "f32":
        ob = torch.full((B, H, W, ostride), 7.0, dtype=torch.float32, device="cuda")
        a.out_f32 = ob.data_ptr() + out_off * 4
        a.out32_mode = L.OUT32_NHWC
        a.out32_pix_stride = ostride
    elif out_kind == "atomic":
        ob = torch.zeros((B, H, W, ostride), dtype=torch.float32, device="cuda")
        a.out_f32 = ob.data_ptr() + out_off * 4
        a.out32_mode = L.OUT32_NHWC_ATOMIC
        a.out32_pix_stride = ostride
    elif out_kind == "nchw":
        ob = torch.full((B, outc, H, W), 7.0, dtype=torch.float32, device="cuda")
        a.out_f32 = ob.data_ptr()
        a.out32_mode = L.OUT32_NCHW
    a.mt, a.n_tile, a.splits = mt, n_tile, splits
    L.check(lib.ssr_conv_tc(C.byref(a), None))
    torch.cuda.synchronize()
    if out_kind == "nchw":
        return ob.cpu(), None
    full = ob.float().cpu()
    out = full[..., out_off:out_off + outc].permute(0, 3, 1, 2).contiguous()
    return out, full


def rel_err(got, ref):
    scale = ref.abs().max().item() + 1e-12
    return (got - ref).abs().max().item() / scale


def make(B, Cin, Cout, H, W, r=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(B, Cin, H, W, generator=g))
    w = bf16_round(torch.randn(Cout, Cin, r, r, generator=g) / (Cin * r * r) ** 0.5)
    return x, w


# ------------------------------------------------------------------ R = 1 (plain GEMM)
@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (256, 64, 32), (384, 256, 128), (256, 96, 64), (1000, 192, 48)])
def test_gemm_1x1(M, K, N):
    x, w = make(1, K, N, 1, M, r=1, seed=M + K + N)
    got, _ = run_conv(x, w, out_kind="f32")
    ref = F.conv2d(x, w)
    assert rel_err(got, ref) < 1e-4


# ------------------------------------------------------------------ 3x3, the generator's shapes
@pytest.mark.parametrize("B,Cin,Cout,H,W", [
    (1, 64, 32, 32, 32),      # RDB conv1
    (2, 96, 32, 32, 32),      # RDB conv2 (ragged last chunk: 2 k-slices)
    (2, 192, 64, 32, 32),     # RDB conv5
    (1, 64, 64, 64, 64),      # conv_up1
    (1, 64, 64, 128, 128),    # conv_up2 / conv_hr
    (2, 128, 64, 16, 16),     # discriminator inner scale
    (1, 32, 64, 32, 32),      # conv_first with a channel-padded input
    (1, 64, 32, 24, 24),      # ragged tile (TW=24, TH=5)
    (1, 64, 32, 40, 136),     # width > 128: two x tiles, the second clipped
])
def test_conv3x3_plain(B, Cin, Cout, H, W):
    x, w = make(B, Cin, Cout, H, W, seed=Cin + Cout + H)
    got, _ = run_conv(x, w, out_kind="f32")
    ref = F.conv2d(x, w, padding=1)
    assert rel_err(got, ref) < 1e-4


def test_conv3x3_bf16_out_bias_lrelu():
    x, w = make(2, 128, 32, 32, 32, seed=5)
    b = torch.randn(32)
    got, _ = run_conv(x, w, bias=b, act=1)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.2)
    assert rel_err(got, ref) < 2 ** -8


def test_conv3x3_mt2():
    x, w = make(2, 128, 32, 32, 32, seed=6)
    got, _ = run_conv(x, w, out_kind="f32", mt=2)
    ref = F.conv2d(x, w, padding=1)
    assert rel_err(got, ref) < 1e-4


def test_conv3x3_n_tiles():
    # Cout = 256 split over two CTAs in N (n_tile = 128)
    x, w = make(1, 128, 256, 16, 16, seed=7)
    got, _ = run_conv(x, w, out_kind="f32")
    ref = F.conv2d(x, w, padding=1)
    assert rel_err(got, ref) < 1e-4


def test_dense_block_slices():
    """conv reads channels [0,160) of a 192-wide buffer and writes its 32 outputs into [160,192)."""
    g = torch.Generator().manual_seed(8)
    x = bf16_round(torch.randn(2, 160, 32, 32, generator=g))
    w = bf16_round(torch.randn(32, 160, 3, 3, generator=g) / 38.0)
    b = torch.randn(32, generator=g)
    got, full = run_conv(x, w, bias=b, act=1, in_stride=192, out_stride=192, out_off=160)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.2)
    assert rel_err(got, ref) < 2 ** -8
    # the rest of the output buffer is untouched
    assert torch.all(full[..., :160] == 7.0)


def test_residual_epilogue():
    """x5*0.2 + x and the RRDB-level (x5*0.2 + x)*0.2 + x0 in one epilogue."""
    g = torch.Generator().manual_seed(9)
    x, w = make(1, 192, 64, 32, 32, seed=9)
    b = torch.randn(64, generator=g)
    r1 = bf16_round(torch.randn(1, 64, 32, 32, generator=g))
    r2 = bf16_round(torch.randn(1, 64, 32, 32, generator=g))
    got, _ = run_conv(x, w, bias=b, s0=0.04, res1=r1, s1=0.2, res2=r2, s2=1.0, out_kind="f32")
    ref = (F.conv2d(x, w, b, padding=1) * 0.2 + r1) * 0.2 + r2
    assert rel_err(got, ref) < 1e-4
    got, _ = run_conv(x, w, bias=b, s0=0.2, res1=r1, s1=1.0, res1_f32=True, out_kind="f32")
    ref = F.conv2d(x, w, b, padding=1) * 0.2 + r1
    assert rel_err(got, ref) < 1e-4


def test_mask_epilogue():
    """the activation-derivative mask shapes the bf16 output only; the f32 output keeps the unmasked running sum"""
    g = torch.Generator().manual_seed(10)
    x, w = make(1, 64, 96, 32, 32, seed=10)
    act = bf16_round(torch.randn(1, 96, 32, 32, generator=g))
    got, _ = run_conv(x, w, mask=act, mask_lo=64)
    ref = F.conv2d(x, w, padding=1)
    ref[:, 64:] = ref[:, 64:] * torch.where(act[:, 64:] > 0, 1.0, 0.2)
    assert rel_err(got, ref) < 2 ** -8
    got, _ = run_conv(x, w, mask=act, mask_lo=0, mask_relu=1)
    ref = F.conv2d(x, w, padding=1) * (act > 0).float()
    assert rel_err(got, ref) < 2 ** -8
    got, _ = run_conv(x, w, mask=act, mask_lo=0, out_kind="f32")
    assert rel_err(got, F.conv2d(x, w, padding=1)) < 1e-4


def test_small_cout_nchw():
    """conv_last: 64 -> 3 channels written straight to an NCHW f32 image."""
    x, w = make(2, 64, 3, 32, 32, seed=11)
    b = torch.randn(3)
    got, _ = run_conv(x, w, bias=b, out_kind="nchw")
    ref = F.conv2d(x, w, b, padding=1)
    assert rel_err(got, ref) < 1e-4


def test_split_k_atomic():
    x, w = make(1, 256, 64, 32, 32, seed=12)
    b = torch.randn(64)
    got, _ = run_conv(x, w, bias=b, out_kind="atomic", splits=4)
    ref = F.conv2d(x, w, b, padding=1)
    assert rel_err(got, ref) < 1e-4


def test_dgrad_pack():
    """input gradient = same kernel with SSR_PACK_DGRAD weights (flipped taps, swapped channels)."""
    L, _ = _lib()
    g = torch.Generator().manual_seed(13)
    w = bf16_round(torch.randn(32, 96, 3, 3, generator=g) / 29.0)   # conv 96 -> 32
    dy = bf16_round(torch.randn(2, 32, 32, 32, generator=g))
    got, _ = run_conv(dy, w, out_kind="f32", mode=L.PACK_DGRAD)
    ref = F.conv_transpose2d(dy, w, padding=1)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-4


def test_bad_args_fail_loudly():
    L, lib = _lib()
    a = L.ConvTcArgs()
    a.r = 5
    rc = lib.ssr_conv_tc(C.byref(a), None)
    assert rc == -1
    assert b"r must be" in lib.ssr_last_error()
