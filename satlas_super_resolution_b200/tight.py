"""Tight-parity ("split-bf16") forward of SSR_RRDBNet (/root/reference/ssr/archs/rrdbnet_arch.py:116-137).

The production path rounds conv operands to bf16 (2^-8 relative); the reference's own GPU arithmetic is TF32 (2^-11) and its CPU
arithmetic fp32.  To tell KERNEL errors from operand-rounding noise this module evaluates the same network through the same
tcgen05 conv kernel (ssr_conv_tc) with every operand carried as a PAIR of bf16 values whose sum holds 16 mantissa bits:

    a = a_hi + a_lo,  w = w_hi + w_lo,    a * w ~= a_hi * w_hi + a_lo * w_hi + a_hi * w_lo      (the dropped a_lo * w_lo is ~2^-18)

i.e. three bf16 launches per convolution whose f32 accumulators are summed, bias / LeakyReLU / residuals applied and the result
split into the next layer's (hi, lo) pair by one elementwise kernel (ssr_split_finish); the weight residuals w - bf16(w) are
packed by ssr_pack_conv_weight(mode | SSR_PACK_LO).  Relative error per layer ~2^-16 -- tighter than TF32 -- so the output agrees
with the fp32 oracle to ~1e-5 at 23 blocks (tests/test_generator_gpu.py::test_split_bf16_forward_matches_fp32_oracle), where the
bf16 production forward sits at 7e-3.  A validation mode: 4x the launches, no CUDA graph, forward only.
"""
import ctypes as C

import torch

from . import _lib as L
from .ops import Act, conv_args, cur_stream, lib, round_up


class _SplitConv:
    """bf16(w) and w - bf16(w) of one convolution as packed tensor-core operands"""

    def __init__(self, weight, bias, cin_buf, device):
        self.cout, self.cin, self.r, _ = weight.shape
        self.bias = bias
        k_pad = round_up(cin_buf, 64)
        n_pad = C.c_int32(0)
        nbytes = lib().ssr_packed_weight_bytes(k_pad, self.cout, self.r, C.byref(n_pad))
        self.n_pad = n_pad.value
        self.hi = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.lo = torch.empty(nbytes, dtype=torch.uint8, device=device)
        w = weight.detach().to(device, torch.float32).contiguous()
        for dst, mode in ((self.hi, L.PACK_FWD), (self.lo, L.PACK_FWD | L.PACK_LO)):
            L.check(lib().ssr_pack_conv_weight(w.data_ptr(), self.cout, self.cin, self.r, mode, None, dst.data_ptr(), k_pad, self.n_pad,
                                               cur_stream()))
        torch.cuda.current_stream().synchronize()   # `w` may be a temporary


class _Pair:
    """an activation as two NHWC bf16 buffers: value = hi + lo"""

    def __init__(self, B, H, W, Cc, device):
        self.hi, self.lo = Act(B, H, W, Cc, device, zero=True), Act(B, H, W, Cc, device, zero=True)
        self.B, self.H, self.W, self.C = B, H, W, Cc

    @property
    def npix(self):
        return self.B * self.H * self.W


class SplitBf16RRDBNet:
    def __init__(self, params, num_in_ch, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32, device=None):
        """params: the reference state_dict schema (conv_first / body.<i>.rdb<j>.conv<k> / conv_body / conv_up<u> / conv_hr /
        conv_last, .weight / .bias); host tensors are copied to `device` (default: the current CUDA device)"""
        if scale not in (1, 2, 4, 8, 16):
            raise ValueError(f"scale {scale} (rrdbnet_arch.py:92-109 builds 1, 2, 4, 8, 16)")
        pdev = next(iter(params.values())).device
        self.device = torch.device(device) if device is not None else (pdev if pdev.type == "cuda" else torch.device("cuda", torch.cuda.current_device()))
        if self.device.type != "cuda":
            raise RuntimeError("SplitBf16RRDBNet: CUDA only (there is no CPU path)")
        self.cin, self.cout, self.scale = num_in_ch, num_out_ch, scale
        self.nf, self.nb, self.g = num_feat, num_block, num_grow_ch
        self.unshuffle = {1: 4, 2: 2}.get(scale, 1)
        self.cin_eff = num_in_ch * self.unshuffle ** 2
        self.cin_pad = round_up(self.cin_eff, 16)
        self.n_up = {1: 2, 2: 2, 4: 2, 8: 3, 16: 4}[scale]
        nf, g = self.nf, self.g
        cv = {}

        def mk(name, cin_buf):
            cv[name] = _SplitConv(params[f"{name}.weight"], params[f"{name}.bias"].detach().to(self.device, torch.float32).contiguous(),
                                  cin_buf, self.device)

        mk("conv_first", self.cin_pad)
        for i in range(num_block):
            for j in (1, 2, 3):
                for k in range(1, 6):
                    mk(f"body.{i}.rdb{j}.conv{k}", nf + (k - 1) * g)
        mk("conv_body", nf)
        for u in range(1, self.n_up + 1):
            mk(f"conv_up{u}", nf)
        mk("conv_hr", nf)
        mk("conv_last", nf)
        self.cv = cv

    # one convolution: three bf16 launches into f32 scratch, then the epilogue kernel
    def _conv(self, name, src, cin, s, act=0, s0=1.0, r1=None, w1=0.0, r2=None, w2=0.0, out_f32=None, dst=None, dst_ch=0, c_pad=None):
        c = self.cv[name]
        B, H, W = src.B, src.H, src.W
        cs = round_up(c.cout, 16)
        sums = [torch.empty((src.npix, cs), dtype=torch.float32, device=self.device) for _ in range(3)]
        for buf, (x, wp) in zip(sums, ((src.hi, c.hi), (src.lo, c.hi), (src.hi, c.lo))):
            a = conv_args(x.ptr(), B, H, W, x.stride, cin, wp.data_ptr(), c.r, c.cout, c.n_pad,
                          out32=buf.data_ptr(), out32_mode=L.OUT32_NHWC, out32_stride=cs)
            L.check(lib().ssr_conv_tc(C.byref(a), s))
        L.check(lib().ssr_split_finish(sums[0].data_ptr(), sums[1].data_ptr(), sums[2].data_ptr(), cs, src.npix, c.cout,
                                       c.bias.data_ptr(), act, s0, r1.data_ptr() if r1 is not None else None, w1,
                                       r2.data_ptr() if r2 is not None else None, w2,
                                       out_f32.data_ptr() if out_f32 is not None else None,
                                       dst.hi.ptr(dst_ch) if dst is not None else None, dst.lo.ptr(dst_ch) if dst is not None else None,
                                       dst.C if dst is not None else 0, c_pad if c_pad is not None else c.cout, s))

    @torch.no_grad()
    def forward(self, x):
        """x: f32 NCHW cuda tensor [B, num_in_ch, h, w] -> f32 NCHW [B, num_out_ch, scale*h, scale*w]"""
        assert x.is_cuda and x.dtype == torch.float32
        f = self.unshuffle
        if f > 1:   # pixel_unshuffle (arch_util.py:769-785): a pure re-indexing
            b, c_, hh, ww = x.shape
            x = x.view(b, c_, hh // f, f, ww // f, f).permute(0, 1, 3, 5, 2, 4).reshape(b, c_ * f * f, hh // f, ww // f)
        B, Cc, h, w = x.shape
        assert Cc == self.cin_eff
        dev, nf, g = self.device, self.nf, self.g
        cw = nf + 4 * g
        s = cur_stream()
        lb = lib()
        npix = B * h * w
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        in0 = _Pair(B, h, w, self.cin_pad, dev)
        L.check(lb.ssr_split_finish(x_nhwc.data_ptr(), None, None, Cc, npix, Cc, None, 0, 1.0, None, 0.0, None, 0.0, None,
                                    in0.hi.ptr(), in0.lo.ptr(), self.cin_pad, self.cin_pad, s))
        n_rdb = 3 * self.nb
        bufs = [_Pair(B, h, w, cw, dev) for _ in range(min(n_rdb, 2))]
        trunk = [torch.empty((npix, nf), dtype=torch.float32, device=dev) for _ in range(n_rdb + 1)]
        body_out = _Pair(B, h, w, nf, dev)
        self._conv("conv_first", in0, self.cin_pad, s, out_f32=trunk[0], dst=bufs[0], dst_ch=0, c_pad=nf)
        for i in range(n_rdb):
            blk, j = divmod(i, 3)
            cur = bufs[i % 2]
            nxt, nxt_pad = (bufs[(i + 1) % 2], nf) if i + 1 < n_rdb else (body_out, nf)
            for k in range(1, 5):       # x_k = lrelu(conv_k(cat(x, x1..x_{k-1}))) -- rrdbnet_arch.py:39-42
                cin = nf + (k - 1) * g
                self._conv(f"body.{blk}.rdb{j + 1}.conv{k}", cur, cin, s, act=1, dst=cur, dst_ch=cin, c_pad=g)
            if j < 2:                   # x5 * 0.2 + x -- :43-44
                self._conv(f"body.{blk}.rdb{j + 1}.conv5", cur, cw, s, s0=0.2, r1=trunk[i], w1=1.0, out_f32=trunk[i + 1],
                           dst=nxt, dst_ch=0, c_pad=nxt_pad)
            else:                       # (x5 * 0.2 + x_rdb3) * 0.2 + x_rrdb -- :68
                self._conv(f"body.{blk}.rdb{j + 1}.conv5", cur, cw, s, s0=0.04, r1=trunk[i], w1=0.2, r2=trunk[3 * blk], w2=1.0,
                           out_f32=trunk[i + 1], dst=nxt, dst_ch=0, c_pad=nxt_pad)
        feat = _Pair(B, h, w, nf, dev)
        self._conv("conv_body", body_out, nf, s, r1=trunk[0], w1=1.0, dst=feat, c_pad=nf)      # feat = feat + body_feat -- :124-125
        src, hh, ww = feat, h, w
        for u in range(self.n_up):     # conv_up(F.interpolate(feat, scale_factor=2, mode='nearest')) -- :127-134
            up = _Pair(B, hh * 2, ww * 2, nf, dev)
            for a, b_ in ((src.hi, up.hi), (src.lo, up.lo)):
                L.check(lb.ssr_upsample_nearest(a.ptr(), nf, b_.ptr(), nf, B, hh, ww, nf, 2, s))
            hh, ww = hh * 2, ww * 2
            out = _Pair(B, hh, ww, nf, dev)
            self._conv(f"conv_up{u + 1}", up, nf, s, act=1, dst=out, c_pad=nf)
            src = out
        hr = _Pair(B, hh, ww, nf, dev)
        self._conv("conv_hr", src, nf, s, act=1, dst=hr, c_pad=nf)                              # :136
        out = torch.empty((B * hh * ww, self.cout), dtype=torch.float32, device=dev)
        self._conv("conv_last", hr, nf, s, out_f32=out)
        return out.view(B, hh, ww, self.cout).permute(0, 3, 1, 2).contiguous()
