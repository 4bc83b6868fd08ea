"""GPU parity of the tcgen05 implicit-GEMM conv (ssr_conv_tc through the C ABI) against torch CPU fp32.

Oracle = F.conv2d in fp32 on the CPU fed the SAME bf16-rounded activations and weights, so the only
differences left are fp32 accumulation order and the final bf16 rounding of the output:
tolerance = 2^-8 relative to the output scale for bf16 outputs, 1e-4 for f32 outputs.
"""
import ctypes as C
import os

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


@pytest.mark.parametrize("B,Cin,Cout,H,W,mt", [
    (16, 64, 64, 64, 64, 0),     # one chunk, one M tile per CTA (512 tiles)
    (20, 64, 64, 64, 64, 2),     # two stacked M tiles
    (32, 64, 64, 128, 64, 0),    # enough tiles for the automatic two-M-tile choice
    (16, 128, 64, 64, 44, 0),    # two chunks; ragged last strip (44 = 5 * 8 + 4)
    (20, 64, 96, 48, 64, 2),     # 96 output channels; 48 rows: the second M tile of the last tile row is half outside
    (20, 192, 32, 40, 40, 0),    # three chunks, narrow N
    (20, 16, 64, 64, 64, 0),     # a 16-channel input (conv9^T, VGG conv1_1): one K step per tap
    (20, 32, 64, 64, 64, 2),     # the discriminator's conv0 (27 -> 32 padded channels)
    (16, 96, 32, 64, 64, 0),     # short second chunk (two K steps)
])
def test_conv3x3_halo_tile_stationary_weights(B, Cin, Cout, H, W, mt):
    """the 8-pixel-strip form (halo in shared memory, 9 taps per stage, stationary weights) against torch, through the short
    epilogue (bias + LeakyReLU -> bf16; mask + bf16 residual) and the general one (f32 output)"""
    L, lib = _lib()
    x, w = make(B, Cin, Cout, H, W, seed=Cin + Cout + W)
    g = torch.Generator().manual_seed(3)
    b = torch.randn(Cout, generator=g)
    n_halo, n_lean = lib.ssr_debug_conv_path_count(0), lib.ssr_debug_conv_path_count(1)
    ref = F.conv2d(x, w, b, padding=1)
    got, _ = run_conv(x, w, bias=b, act=1, mt=mt)
    assert rel_err(got, F.leaky_relu(ref, 0.2)) < 2 ** -8
    act = bf16_round(torch.randn(B, Cout, H, W, generator=g))
    r1 = bf16_round(torch.randn(B, Cout, H, W, generator=g))
    got, _ = run_conv(x, w, bias=b, s0=0.5, res1=r1, s1=2.0, mask=act, mt=mt)
    assert rel_err(got, (ref * 0.5 + 2.0 * r1) * torch.where(act > 0, 1.0, 0.2)) < 2 ** -7
    got, _ = run_conv(x, w, bias=b, out_kind="f32", mt=mt)
    assert rel_err(got, ref) < 1e-4
    if os.environ.get("SSR_CONV_HALO", "1") != "0":
        assert lib.ssr_debug_conv_path_count(0) - n_halo == 3
    if os.environ.get("SSR_CONV_LEAN", "1") != "0":
        assert lib.ssr_debug_conv_path_count(1) - n_lean == 2


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


@pytest.mark.parametrize("B,H,W,planar", [(2, 32, 32, False), (32, 32, 32, False), (3, 24, 40, False),
                                            (2, 32, 32, True), (32, 32, 32, True), (3, 32, 64, True), (5, 32, 16, True)])
def test_chain_equals_plain_launches(B, H, W, planar):
    """ssr_conv_tc_chain (five dense-block convs in ONE launch) must be bit-identical to the same five ssr_conv_tc calls; run
    three times (barrier phases / counters must re-arm).  planar (the layout the generator uses for its f32 trunk): 32-row images
    take the shared-memory-resident kernel (8-pixel strips, halo columns through DSMEM, one cluster per image: 2 to 8 CTAs);
    NHWC f32 operands: the cluster-synchronised chain over global memory (32 x 32) or plain launches."""
    L, lib = _lib()
    f32_kind, f32_mode = (L.SSR_F32_PLANAR4, L.OUT32_PLANAR4) if planar else (L.SSR_F32, L.OUT32_NHWC)
    nf, g = 64, 32
    cw = nf + 4 * g
    torch.manual_seed(B * 1000 + H)
    x0 = torch.randn(B, H, W, nf) * 0.5
    ws, packs, biases = [], [], []
    for k in range(5):
        cin, cout = nf + k * g, (g if k < 4 else nf)
        wk = torch.randn(cout, cin, 3, 3) * (1.0 / (3.0 * cin ** 0.5))
        packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD, k_pad=(cin + 63) // 64 * 64)
        packs.append((packed, n_pad, cin, cout))
        biases.append((torch.randn(cout) * 0.1).cuda())
    trunk = torch.randn(B, H, W, nf).cuda()

    def make(buf, nxt, t32):
        arr = (L.ConvTcArgs * 5)()
        for k in range(5):
            packed, n_pad, cin, cout = packs[k]
            a = arr[k]
            a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cin = buf.data_ptr(), B, H, W, cw, cin
            a.w_packed, a.r, a.cout, a.n_pad = packed.data_ptr(), 3, cout, n_pad
            a.bias = biases[k].data_ptr()
            if k < 4:
                a.act, a.s0 = 1, 1.0
                a.out_bf16, a.out_pix_stride = buf.data_ptr() + 2 * cin, cw
            else:
                a.act, a.s0 = 0, 0.2
                a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = trunk.data_ptr(), f32_kind, nf, 1.0
                a.out_bf16, a.out_pix_stride = nxt.data_ptr(), nf
                a.out_f32, a.out32_mode, a.out32_pix_stride = t32.data_ptr(), f32_mode, nf
        return arr

    def fresh():
        buf = torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda")
        buf[..., :nf] = x0.cuda().to(torch.bfloat16)
        return buf, torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda"), torch.zeros(B, H, W, nf, device="cuda")

    s = torch.cuda.current_stream().cuda_stream
    buf_a, nxt_a, t_a = fresh()
    arr = make(buf_a, nxt_a, t_a)
    for k in range(5):
        L.check(lib.ssr_conv_tc(C.byref(arr[k]), s))
    torch.cuda.synchronize()
    assert nxt_a.float().abs().max() > 0
    for _ in range(3):
        buf_b, nxt_b, t_b = fresh()
        arr_b = make(buf_b, nxt_b, t_b)
        n0, r0 = lib.ssr_launch_count(), lib.ssr_debug_resident_launches()
        L.check(lib.ssr_conv_tc_chain(arr_b, 5, s))
        torch.cuda.synchronize()
        # 32 x 32 images are 4 (or 8) tiles: one cluster per image, a single launch; 24 x 40 is 12 tiles: plain launches
        resident = planar and os.environ.get("SSR_CONV_RESIDENT", "1") != "0"
        # without the resident kernel: images of <= 8 pixel tiles (32 x 32, 32 x 16) chain inside one cluster launch, others are plain launches
        assert lib.ssr_launch_count() - n0 == (1 if (H, W) in ((32, 32), (32, 16)) or resident else 5)
        assert lib.ssr_debug_resident_launches() - r0 == (1 if resident else 0)
        assert torch.equal(buf_a, buf_b)
        assert torch.equal(nxt_a, nxt_b)
        assert torch.equal(t_a, t_b)


@pytest.mark.parametrize("B", [2, 32])
def test_chain_wide_layers_in_place_gradient(B):
    """the input-gradient chain of a dense block: N = 192, 160, 128, 96, 64 output channels (several N tiles per CTA),
    a running f32 gradient read and rewritten in place, derivative masks -- chained launch == five plain launches, bit for bit"""
    L, lib = _lib()
    H = W = 32
    nf, g = 64, 32
    cw = nf + 4 * g
    torch.manual_seed(7 + B)
    xin = (torch.randn(B, H, W, nf) * 0.5).cuda().to(torch.bfloat16)
    cur = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)          # forward activations -> LeakyReLU derivative mask
    g32_0 = torch.randn(B, H, W, cw).cuda() * 0.1
    layers = []
    for k in range(5, 0, -1):
        nk = nf + (k - 1) * g                      # output channels of this layer; it reads the g-channel dY slot at nk
        cin = nf if k == 5 else g
        wk = torch.randn(nk, cin, 3, 3) * (1.0 / (3.0 * cin ** 0.5))
        packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD, k_pad=64)
        layers.append((k, nk, cin, packed, n_pad))

    def make(dg, G32, gout, bg, acc=False):
        arr = (L.ConvTcArgs * 5)()
        for i, (k, nk, cin, packed, n_pad) in enumerate(layers):
            a = arr[i]
            a.x = xin.data_ptr() if k == 5 else dg.data_ptr() + 2 * nk
            a.x_pix_stride = nf if k == 5 else cw
            a.n_img, a.h, a.w, a.cin = B, H, W, cin
            a.w_packed, a.r, a.cout, a.n_pad = packed.data_ptr(), 3, nk, n_pad
            a.s0 = 0.2 if k == 5 else 1.0
            a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = G32.data_ptr(), L.SSR_F32_PLANAR4, cw, 1.0
            if k == 5:
                a.res1_cmax = nf
            # acc: channels without a bf16 output are added into the running sum with vector reductions (same values)
            a.out_f32, a.out32_mode, a.out32_pix_stride = G32.data_ptr(), (L.OUT32_PLANAR4_ACC if acc and k > 1 else L.OUT32_PLANAR4), cw
            # only the top g-channel slot is stored as (masked) bf16; its pixel sums go to row i of bg, scaled
            a.bias_grad, a.bias_grad_scale = bg.data_ptr() + 4 * 64 * i, 0.5
            if k > 1:
                a.mask, a.mask_pix_stride, a.mask_lo, a.out_lo = cur.data_ptr(), cw, nk - g, nk - g
                a.out_bf16, a.out_pix_stride = dg.data_ptr(), cw
            else:
                a.out_bf16, a.out_pix_stride = gout.data_ptr(), nf
        return arr

    def fresh():
        return (torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda"), g32_0.clone(),
                torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda"), torch.zeros(5, 64, device="cuda"))

    s = torch.cuda.current_stream().cuda_stream
    ref = fresh()
    arr = make(*ref)
    for i in range(5):
        L.check(lib.ssr_conv_tc(C.byref(arr[i]), s))
    torch.cuda.synchronize()
    assert ref[2].float().abs().max() > 0
    # the fused bias gradient = scale * pixel sum of what the bf16 output received (f32 values before the bf16 rounding)
    for i, (k, nk, cin, packed, n_pad) in enumerate(layers):
        slot = ref[0][..., nk - g:nk] if k > 1 else ref[2]
        want = 0.5 * slot.float().sum(dim=(0, 1, 2))
        got_bg = ref[3][i, :want.numel()]
        assert torch.allclose(got_bg, want, rtol=2e-2, atol=2e-2 * want.abs().max().item()), (k, (got_bg - want).abs().max())
    assert ref[0][..., :nf].float().abs().max() == 0   # below every layer's out_lo: never stored as bf16
    for _ in range(3):
        got = fresh()
        arr_b = make(*got, acc=(_ > 0))     # first pass: same modes as the reference; then the reduction form
        n0 = lib.ssr_launch_count()
        L.check(lib.ssr_conv_tc_chain(arr_b, 5, s))
        torch.cuda.synchronize()
        assert lib.ssr_launch_count() - n0 == 1
        for r, o in zip(ref[:3], got[:3]):
            assert torch.equal(r, o)
        assert torch.allclose(ref[3], got[3], rtol=1e-4, atol=1e-4 * ref[3].abs().max().item())   # atomics: order differs


@pytest.mark.parametrize("B,W", [(2, 32), (32, 32), (3, 64), (5, 16)])
def test_chain_acc_running_sum_in_tensor_memory(B, W):
    """ssr_conv_tc_chain_acc: the five input-gradient layers add into ONE accumulator that never leaves tensor memory; each layer
    emits only its top 32-channel slot, the last one the 64 block-input channels plus the incoming gradient.  Reference = the
    same layers as plain launches accumulating through an f32 buffer in global memory (f32 summation order differs: tolerance)."""
    L, lib = _lib()
    H = 32
    nf, g = 64, 32
    cw = nf + 4 * g
    P = B * H * W
    if not lib.ssr_conv_tc_chain_acc_supported(B, H, W, cw):
        assert W == 64 and os.environ.get("SSR_CONV_RESIDENT") == "0"   # 16 tiles per image: only the resident kernel takes it
        pytest.skip("32 x 64 images need the shared-memory-resident kernel")
    assert lib.ssr_conv_tc_chain_acc_supported(B, 128, 128, cw) == 0
    torch.manual_seed(11 + B)
    xin = (torch.randn(B, H, W, nf) * 0.5).cuda().to(torch.bfloat16)
    cur = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)
    incoming = (torch.randn(nf // 4, P, 4) * 0.3).cuda()          # planar f32 gradient arriving from the block above
    layers = []
    for k in range(5, 0, -1):
        nk = nf + (k - 1) * g
        cin = nf if k == 5 else g
        wk = torch.randn(nk, cin, 3, 3) * (1.0 / (3.0 * cin ** 0.5))
        packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD, k_pad=64)
        layers.append((k, nk, cin, packed, n_pad))

    def make(dg, G32, gout, bg, acc):
        arr = (L.ConvTcArgs * 5)()
        for i, (k, nk, cin, packed, n_pad) in enumerate(layers):
            a = arr[i]
            a.x = xin.data_ptr() if k == 5 else dg.data_ptr() + 2 * nk
            a.x_pix_stride = nf if k == 5 else cw
            a.n_img, a.h, a.w, a.cin = B, H, W, cin
            a.w_packed, a.r, a.cout, a.n_pad, a.s0 = packed.data_ptr(), 3, nk, n_pad, 1.0
            a.bias_grad, a.bias_grad_scale = bg.data_ptr() + 4 * 64 * i, 0.5
            if k > 1:
                a.mask, a.mask_pix_stride, a.mask_lo, a.out_lo = cur.data_ptr(), cw, nk - g, nk - g
                a.out_bf16, a.out_pix_stride = dg.data_ptr(), cw
            else:
                a.out_bf16, a.out_pix_stride = gout.data_ptr(), nf
            if acc:
                if k == 1:   # only the last layer touches global f32 memory: sum + incoming -> out32
                    a.res1, a.res1_kind, a.s1 = incoming.data_ptr(), L.SSR_F32_PLANAR4, 1.0
                    a.out_f32, a.out32_mode = G32.data_ptr(), L.OUT32_PLANAR4
            else:
                # plain launches: the running sum lives in G32 (planar, cw channels); it starts as `incoming` in channels [0, nf)
                if k == 5:
                    a.res1, a.res1_kind, a.s1, a.res1_cmax = G32.data_ptr(), L.SSR_F32_PLANAR4, 1.0, nf
                else:
                    a.res1, a.res1_kind, a.s1 = G32.data_ptr(), L.SSR_F32_PLANAR4, 1.0
                a.out_f32, a.out32_mode = G32.data_ptr(), L.OUT32_PLANAR4
        return arr

    def fresh():
        G32 = torch.zeros(cw // 4, P, 4, device="cuda")
        return (torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda"), G32,
                torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda"), torch.zeros(5, 64, device="cuda"))

    s = torch.cuda.current_stream().cuda_stream
    ref = fresh()
    ref[1][:nf // 4].copy_(incoming)
    arr = make(*ref, acc=False)
    for i in range(5):
        L.check(lib.ssr_conv_tc(C.byref(arr[i]), s))
    torch.cuda.synchronize()
    for _ in range(2):
        got = fresh()
        arr_b = make(*got, acc=True)
        n0, r0 = lib.ssr_launch_count(), lib.ssr_debug_resident_launches()
        L.check(lib.ssr_conv_tc_chain_acc(arr_b, 5, s))
        torch.cuda.synchronize()
        assert lib.ssr_launch_count() - n0 == 1
        # 32-row images: the dY slots also stay in shared memory (resident tile), not only the running sum in tensor memory
        assert lib.ssr_debug_resident_launches() - r0 == (1 if os.environ.get("SSR_CONV_RESIDENT", "1") != "0" else 0)
        scale = ref[0][..., nf:].float().abs().max().item()
        assert (ref[0][..., nf:].float() - got[0][..., nf:].float()).abs().max().item() <= 2 ** -7 * scale      # the four dY slots
        assert rel_err(got[2].float().cpu(), ref[2].float().cpu()) < 2 ** -7                                      # block-input gradient, bf16
        # ... and f32: a dY slot that rounds to the other bf16 neighbour (different f32 summation order) moves later layers slightly
        assert rel_err(got[1][:nf // 4].cpu(), ref[1][:nf // 4].cpu()) < 5e-3   # measured 4e-4 at B = 2
        assert got[1][nf // 4:].abs().max() == 0                                                                    # nothing else leaves TMEM
        assert torch.allclose(ref[3], got[3], rtol=1e-3, atol=1e-3 * ref[3].abs().max().item())                    # bias gradients


@pytest.mark.parametrize("B,W,nblk", [(2, 32, 3), (32, 32, 3), (3, 64, 2), (5, 16, 3), (2, 32, 12), (3, 64, 7)])
def test_several_dense_blocks_in_one_resident_launch(B, W, nblk):
    """ssr_conv_tc_chain with 5 * nblk layers: consecutive ResidualDenseBlocks (an RRDB: the third block also adds the RRDB-level
    residual) in ONE shared-memory-resident launch -- each block's conv5 hands its 64 channels to the next block through the tile.
    Must be bit-identical to one resident launch per block (the global stores are the same; only where the next block READS
    its input differs).  Run three times: barrier phases re-arm."""
    L, lib = _lib()
    if lib.ssr_rdb_resident_max_blocks(B, 32, W) < nblk:
        pytest.skip("resident multi-block launches are switched off (SSR_CONV_RESIDENT=0 / SSR_RDB_FUSE)")
    H, nf, g = 32, 64, 32
    cw = nf + 4 * g
    P = B * H * W
    torch.manual_seed(B * 100 + W)
    x0 = torch.randn(B, H, W, nf) * 0.5
    trunk0 = (torch.randn(nf // 4, P, 4) * 0.5).cuda()
    packs, biases = [], []
    for b in range(nblk):
        for k in range(5):
            cin, cout = nf + k * g, (g if k < 4 else nf)
            wk = torch.randn(cout, cin, 3, 3) * (1.0 / (3.0 * cin ** 0.5))
            packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD, k_pad=(cin + 63) // 64 * 64)
            packs.append((packed, n_pad, cin, cout))
            biases.append((torch.randn(cout) * 0.1).cuda())

    def make(bufs, out, trunks):
        arr = (L.ConvTcArgs * (5 * nblk))()
        for b in range(nblk):
            buf = bufs[b]
            for k in range(5):
                packed, n_pad, cin, cout = packs[5 * b + k]
                a = arr[5 * b + k]
                a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cin = buf.data_ptr(), B, H, W, cw, cin
                a.w_packed, a.r, a.cout, a.n_pad = packed.data_ptr(), 3, cout, n_pad
                a.bias = biases[5 * b + k].data_ptr()
                if k < 4:
                    a.act, a.s0 = 1, 1.0
                    a.out_bf16, a.out_pix_stride = buf.data_ptr() + 2 * cin, cw
                else:
                    last = b % 3 == 2 or b + 1 == nblk      # every third block closes an RRDB
                    a.act, a.s0 = 0, (0.04 if last else 0.2)
                    a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = trunks[b].data_ptr(), L.SSR_F32_PLANAR4, nf, (0.2 if last else 1.0)
                    if last:   # (x5*0.2 + x_rdb)*0.2 + x_rrdb -- rrdbnet_arch.py:68
                        a.res2, a.res2_kind, a.res2_pix_stride, a.s2 = trunks[b - b % 3].data_ptr(), L.SSR_F32_PLANAR4, nf, 1.0
                    nxt, stride = (out, nf) if b + 1 == nblk else (bufs[b + 1], cw)
                    a.out_bf16, a.out_pix_stride = nxt.data_ptr(), stride
                    a.out_f32, a.out32_mode, a.out32_pix_stride = trunks[b + 1].data_ptr(), L.OUT32_PLANAR4, nf
        return arr

    def fresh():
        bufs = [torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda") for _ in range(nblk)]
        bufs[0][..., :nf] = x0.cuda().to(torch.bfloat16)
        trunks = [trunk0.clone()] + [torch.zeros(nf // 4, P, 4, device="cuda") for _ in range(nblk)]
        return bufs, torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda"), trunks

    s = torch.cuda.current_stream().cuda_stream
    ref = fresh()
    arr = make(*ref)
    for b in range(nblk):      # one resident launch per block
        one = (L.ConvTcArgs * 5)(*[arr[5 * b + k] for k in range(5)])
        L.check(lib.ssr_conv_tc_chain(one, 5, s))
    torch.cuda.synchronize()
    assert ref[1].float().abs().max() > 0
    for _ in range(3):
        got = fresh()
        arr_b = make(*got)
        n0, r0 = lib.ssr_launch_count(), lib.ssr_debug_resident_launches()
        L.check(lib.ssr_conv_tc_chain(arr_b, 5 * nblk, s))
        torch.cuda.synchronize()
        assert lib.ssr_launch_count() - n0 == 1 and lib.ssr_debug_resident_launches() - r0 == 1
        for b in range(nblk):
            assert torch.equal(ref[0][b], got[0][b]), b
            assert torch.equal(ref[2][b + 1], got[2][b + 1]), b
        assert torch.equal(ref[1], got[1])


@pytest.mark.parametrize("B,W,nblk", [(2, 32, 3), (32, 32, 3), (3, 64, 2), (5, 16, 3), (2, 32, 12), (3, 64, 7)])
def test_several_input_gradient_chains_in_one_resident_launch(B, W, nblk):
    """ssr_conv_tc_chain_acc with 5 * nblk layers: the input-gradient chains of consecutive dense blocks in ONE launch; a block's
    64-channel result (sum in tensor memory + incoming f32 gradient) is the next block's incoming bf16 gradient, handed over through
    the tile.  Bit-identical to one launch per block (bias gradients: atomics, tolerance)."""
    L, lib = _lib()
    if lib.ssr_rdb_resident_max_blocks(B, 32, W) < nblk:
        pytest.skip("resident multi-block launches are switched off (SSR_CONV_RESIDENT=0 / SSR_RDB_FUSE)")
    H, nf, g = 32, 64, 32
    cw = nf + 4 * g
    P = B * H * W
    torch.manual_seed(B * 10 + W)
    xin = (torch.randn(B, H, W, nf) * 0.5).cuda().to(torch.bfloat16)
    curs = [torch.randn(B, H, W, cw).cuda().to(torch.bfloat16) for _ in range(nblk)]
    g32_0 = (torch.randn(nf // 4, P, 4) * 0.3).cuda()
    layers = []
    for b in range(nblk):
        for k in range(5, 0, -1):
            nk = nf + (k - 1) * g
            cin = nf if k == 5 else g
            wk = torch.randn(nk, cin, 3, 3) * (1.0 / (3.0 * cin ** 0.5))
            packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD, k_pad=64)
            layers.append((k, nk, cin, packed, n_pad))

    def make(dgs, G32, gouts, bg):
        arr = (L.ConvTcArgs * (5 * nblk))()
        for b in range(nblk):
            dg = dgs[b]
            for i in range(5):
                k, nk, cin, packed, n_pad = layers[5 * b + i]
                a = arr[5 * b + i]
                if k == 5:
                    a.x, a.x_pix_stride = (xin if b == 0 else gouts[b - 1]).data_ptr(), nf
                else:
                    a.x, a.x_pix_stride = dg.data_ptr() + 2 * nk, cw
                a.n_img, a.h, a.w, a.cin = B, H, W, cin
                a.w_packed, a.r, a.cout, a.n_pad, a.s0 = packed.data_ptr(), 3, nk, n_pad, 1.0
                a.bias_grad, a.bias_grad_scale = bg.data_ptr() + 4 * 64 * (5 * b + i), 0.5
                if k > 1:
                    a.mask, a.mask_pix_stride, a.mask_lo, a.out_lo = curs[b].data_ptr(), cw, nk - g, nk - g
                    a.out_bf16, a.out_pix_stride = dg.data_ptr(), cw
                else:
                    a.out_bf16, a.out_pix_stride = gouts[b].data_ptr(), nf
                    a.res1, a.res1_kind, a.s1 = G32.data_ptr(), L.SSR_F32_PLANAR4, 1.0     # running f32 gradient, in place
                    a.out_f32, a.out32_mode = G32.data_ptr(), L.OUT32_PLANAR4
        return arr

    def fresh():
        return ([torch.zeros(B, H, W, cw, dtype=torch.bfloat16, device="cuda") for _ in range(nblk)], g32_0.clone(),
                [torch.zeros(B, H, W, nf, dtype=torch.bfloat16, device="cuda") for _ in range(nblk)],
                torch.zeros(5 * nblk, 64, device="cuda"))

    s = torch.cuda.current_stream().cuda_stream
    ref = fresh()
    arr = make(*ref)
    for b in range(nblk):
        one = (L.ConvTcArgs * 5)(*[arr[5 * b + i] for i in range(5)])
        L.check(lib.ssr_conv_tc_chain_acc(one, 5, s))
    torch.cuda.synchronize()
    assert ref[2][-1].float().abs().max() > 0
    for _ in range(3):
        got = fresh()
        arr_b = make(*got)
        n0, r0 = lib.ssr_launch_count(), lib.ssr_debug_resident_launches()
        L.check(lib.ssr_conv_tc_chain_acc(arr_b, 5 * nblk, s))
        torch.cuda.synchronize()
        assert lib.ssr_launch_count() - n0 == 1 and lib.ssr_debug_resident_launches() - r0 == 1
        for b in range(nblk):
            assert torch.equal(ref[0][b], got[0][b]), b
            assert torch.equal(ref[2][b], got[2][b]), b
        assert torch.equal(ref[1], got[1])
        assert torch.allclose(ref[3], got[3], rtol=1e-3, atol=1e-3 * ref[3].abs().max().item())


@pytest.mark.parametrize("cout", [64, 20])
def test_planar4_f32_operands_match_nhwc(cout):
    """SSR_F32_PLANAR4 residual + SSR_OUT32_PLANAR4 output hold exactly the values of the NHWC f32 forms (layout only)"""
    L, lib = _lib()
    B, H, W, cin = 2, 32, 32, 64
    P = B * H * W
    torch.manual_seed(cout)
    x = (torch.randn(B, H, W, cin) * 0.5).cuda().to(torch.bfloat16)
    wk = torch.randn(cout, cin, 3, 3) * 0.05
    packed, n_pad = pack_weight(L, lib, wk, L.PACK_FWD)
    cq = (cout + 3) // 4 * 4
    res = torch.randn(P, cq).cuda()
    res_planar = res.view(P, cq // 4, 4).permute(1, 0, 2).contiguous()
    s = torch.cuda.current_stream().cuda_stream

    def run(planar):
        a = L.ConvTcArgs()
        a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cin = x.data_ptr(), B, H, W, cin, cin
        a.w_packed, a.r, a.cout, a.n_pad, a.s0 = packed.data_ptr(), 3, cout, n_pad, 0.5
        out = torch.zeros(cq // 4, P, 4, device="cuda") if planar else torch.zeros(P, cq, device="cuda")
        a.res1 = (res_planar if planar else res).data_ptr()
        a.res1_kind, a.res1_pix_stride, a.s1 = (L.SSR_F32_PLANAR4 if planar else L.SSR_F32), cq, 1.5
        a.out_f32, a.out32_mode, a.out32_pix_stride = out.data_ptr(), (L.OUT32_PLANAR4 if planar else L.OUT32_NHWC), cq
        L.check(lib.ssr_conv_tc(C.byref(a), s))
        torch.cuda.synchronize()
        return out.permute(1, 0, 2).reshape(P, cq) if planar else out

    nhwc, planar = run(False), run(True)
    assert nhwc.abs().max() > 0
    assert torch.equal(nhwc[:, :cout], planar[:, :cout])


def test_chains_without_the_resident_kernel_in_subprocess():
    """SSR_CONV_RESIDENT=0: the cluster-synchronised chains over global memory (what shapes that do not fit the resident kernel
    use) must still equal plain launches.  Environment switches are read once per process: run the chain tests in a child."""
    import subprocess
    import sys
    env = dict(os.environ, SSR_CONV_RESIDENT="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", "chain_equals_plain or chain_acc"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 64, 128, 32, 32), (1, 128, 256, 64, 64), (3, 256, 512, 16, 32), (32, 64, 128, 128, 128)])
def test_strided_4x4_conv_forward_and_transposed_parity_classes(B, cin, cout, H, W):
    """discriminator conv1..3 (discriminator_arch.py:30-32, 45-47): the 4 x 4 stride-2 pad-1 conv as an implicit GEMM (TMA element
    strides gather every second pixel per tap; no im2col columns), and its input gradient as four 2 x 2 parity-class convs over dY
    that scatter to (2y + oy, 2x + ox) with the skip-gradient add and the LeakyReLU mask of the producer fused in."""
    from satlas_super_resolution_b200.ops import PackedConv, Packer, conv_args, dgrad_s2_class_ptr
    L, lib = _lib()
    g = torch.Generator().manual_seed(cin + H)
    x = bf16_round(torch.randn(B, cin, H, W, generator=g))
    w = bf16_round(torch.randn(cout, cin, 4, 4, generator=g) / (cin * 16) ** 0.5)
    wd = w.cuda().contiguous()
    cv = PackedConv(wd, None, cin, True, "cuda")
    Packer([cv], "cuda").run(None)
    xb = nhwc_buffer(x)
    out = torch.full((B, H // 2, W // 2, cout), 7.0, dtype=torch.bfloat16, device="cuda")
    a = conv_args(xb.data_ptr(), B, H, W, cin, cin, cv.packed.data_ptr(), 4, cout, cv.n_pad, act=1, out=out.data_ptr(), out_stride=cout, stride=2)
    L.check(lib.ssr_conv_tc(C.byref(a), None))
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(x, w, stride=2, padding=1), 0.2)
    assert rel_err(out.float().cpu().permute(0, 3, 1, 2), ref) < 2 ** -7
    # ---- transposed: dX = (conv_transpose(dY) + skip) * LeakyReLU'(x)
    dy = bf16_round(torch.randn(B, cout, H // 2, W // 2, generator=g))
    skip = bf16_round(torch.randn(B, cin, H, W, generator=g))
    dyb, skb = nhwc_buffer(dy), nhwc_buffer(skip)
    dx = torch.full((B, H, W, cin), 7.0, dtype=torch.bfloat16, device="cuda")
    for cls in range(4):
        oy, ox = divmod(cls, 2)
        a = conv_args(dyb.data_ptr(), B, H // 2, W // 2, cout, cout, dgrad_s2_class_ptr(cv, cls), 2, cin, cv.n_pad_dg,
                      pad_y=1 - oy, pad_x=1 - ox, out_oy=oy, out_ox=ox, mask=xb.data_ptr(), mask_stride=cin, mask_lo=0,
                      res1=skb.data_ptr(), res1_kind=L.SSR_BF16, res1_stride=cin, s1=1.0, out=dx.data_ptr(), out_stride=cin)
        L.check(lib.ssr_conv_tc(C.byref(a), None))
    torch.cuda.synchronize()
    ref_dx = (F.conv_transpose2d(dy, w, stride=2, padding=1) + skip) * torch.where(x > 0, 1.0, 0.2)
    assert rel_err(dx.float().cpu().permute(0, 3, 1, 2), ref_dx) < 2 ** -7
