"""Tight-parity ("split-bf16") evaluation of SSR_RRDBNet (/root/reference/ssr/archs/rrdbnet_arch.py:116-137): forward AND backward.

The production path rounds conv operands to bf16 (2^-8 relative); the reference's own GPU arithmetic is TF32 (2^-11) and its CPU
arithmetic fp32.  To tell KERNEL errors from operand-rounding noise this module evaluates the same network through the same
tcgen05 kernels (ssr_conv_tc for the forward and the input gradients, ssr_wgrad_tc for the weight gradients) with every operand
carried as a PAIR of bf16 values whose sum holds 16 mantissa bits:

    a = a_hi + a_lo,  w = w_hi + w_lo,    a * w ~= a_hi * w_hi + a_lo * w_hi + a_hi * w_lo      (the dropped a_lo * w_lo is ~2^-18)

i.e. three bf16 launches per contraction whose f32 accumulators are summed; bias / LeakyReLU / residuals / derivative masks are
applied and the result is split into the next layer's (hi, lo) pair by one elementwise kernel (ssr_split_finish); weight residuals
w - bf16(w) are packed by ssr_pack_conv_weight(mode | SSR_PACK_LO).  Relative error per layer ~2^-16 -- tighter than TF32 -- so the
forward agrees with the fp32 oracle to ~1e-5 at 23 blocks, where the bf16 production forward sits at 7e-3, and the parameter
gradients agree with autograd of the PLAIN fp32 oracle (its own LeakyReLU pattern) to ~1e-4
(tests/test_generator_gpu.py::test_split_bf16_*).  A validation mode: 4x the launches, eager, no CUDA graph.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _lib as L
from .ops import Act, WgradSet, conv_args, cur_stream, lib, round_up


class _SplitConv:
    """bf16(w) and w - bf16(w) of one convolution as packed tensor-core operands (forward and, optionally, input-gradient form)"""

    def __init__(self, weight, bias, cin_buf, device, want_dgrad=False):
        self.cout, self.cin, self.r, _ = weight.shape
        self.bias = bias
        self.cin_buf = cin_buf
        w = weight.detach().to(device, torch.float32).contiguous()
        n_pad = C.c_int32(0)

        def pack(k_pad, n_out, mode):
            nbytes = lib().ssr_packed_weight_bytes(k_pad, n_out, self.r, C.byref(n_pad))
            pair = []
            for m in (mode, mode | L.PACK_LO):
                dst = torch.empty(nbytes, dtype=torch.uint8, device=device)
                L.check(lib().ssr_pack_conv_weight(w.data_ptr(), self.cout, self.cin, self.r, m, None, dst.data_ptr(), k_pad, n_pad.value,
                                                   cur_stream()))
                pair.append(dst)
            return pair[0], pair[1], n_pad.value

        self.hi, self.lo, self.n_pad = pack(round_up(cin_buf, 64), self.cout, L.PACK_FWD)
        if want_dgrad:
            self.cout_buf = round_up(self.cout, 16)     # channels of the dY buffer the transposed conv reads
            self.dg_hi, self.dg_lo, self.n_pad_dg = pack(round_up(self.cout_buf, 64), self.cin, L.PACK_DGRAD)
        torch.cuda.current_stream().synchronize()   # `w` may be a temporary


class _Pair:
    """an activation (or activation gradient) as two NHWC bf16 buffers: value = hi + lo"""

    def __init__(self, B, H, W, Cc, device):
        self.hi, self.lo = Act(B, H, W, Cc, device, zero=True), Act(B, H, W, Cc, device, zero=True)
        self.B, self.H, self.W, self.C = B, H, W, Cc

    @property
    def npix(self):
        return self.B * self.H * self.W


class SplitBf16RRDBNet:
    def __init__(self, params, num_in_ch, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32, device=None, want_grad=False):
        """params: the reference state_dict schema (conv_first / body.<i>.rdb<j>.conv<k> / conv_body / conv_up<u> / conv_hr /
        conv_last, .weight / .bias); host tensors are copied to `device` (default: the current CUDA device).  want_grad: also build
        the input-gradient operands and f32 gradient buffers (`backward`)."""
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
        self.want_grad = want_grad
        nf, g = self.nf, self.g
        cv = OrderedDict()

        def mk(name, cin_buf):
            cv[name] = _SplitConv(params[f"{name}.weight"], params[f"{name}.bias"].detach().to(self.device, torch.float32).contiguous(),
                                  cin_buf, self.device, want_dgrad=want_grad and name != "conv_first")

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
        self.grads = None
        self._saved = None
        if want_grad:
            self.grads = OrderedDict()
            for name, c in cv.items():
                self.grads[f"{name}.weight"] = torch.zeros((c.cout, c.cin, c.r, c.r), dtype=torch.float32, device=self.device)
                self.grads[f"{name}.bias"] = torch.zeros((c.cout,), dtype=torch.float32, device=self.device)
            self.wg = WgradSet(self.device)
            for name, c in cv.items():
                self.wg.add(name, c, c.cin_buf)
            self.wg.finalize(lambda name: self.grads[f"{name}.weight"])

    # ------------------------------------------------------------------ building blocks
    def _finish(self, sums, npix, c, s, sum_stride=None, bias=None, act=0, s0=1.0, r1=None, r1_stride=0, w1=0.0, r2=None, r2_stride=0, w2=0.0,
                mask=None, mask_stride=0, out_f32=None, out32_stride=0, hi=None, lo=None, out_stride=0, c_pad=None):
        """ssr_split_finish with python-side defaults; sums: 1..3 f32 device pointers (ints)"""
        ps = list(sums) + [None] * (3 - len(sums))
        L.check(lib().ssr_split_finish(ps[0], ps[1], ps[2], sum_stride if sum_stride is not None else c, npix, c, bias, act, s0,
                                       r1, r1_stride, w1, r2, r2_stride, w2, mask, mask_stride, out_f32, out32_stride, hi, lo, out_stride,
                                       c_pad if c_pad is not None else c, s))

    def _three(self, x, wp_hi, wp_lo, B, H, W, cin, cout, n_pad, s, ch=0):
        """a_hi * w_hi, a_lo * w_hi, a_hi * w_lo as three ssr_conv_tc launches -> three NHWC f32 sum buffers [npix, cs]"""
        cs = round_up(cout, 16)
        sums = [torch.empty((B * H * W, cs), dtype=torch.float32, device=self.device) for _ in range(3)]
        for buf, (xb, wp) in zip(sums, ((x.hi, wp_hi), (x.lo, wp_hi), (x.hi, wp_lo))):
            a = conv_args(xb.ptr(ch), B, H, W, xb.stride, cin, wp.data_ptr(), 3, cout, n_pad,
                          out32=buf.data_ptr(), out32_mode=L.OUT32_NHWC, out32_stride=cs)
            L.check(lib().ssr_conv_tc(C.byref(a), s))
        return sums, cs

    def _conv(self, name, src, cin, s, act=0, s0=1.0, r1=None, w1=0.0, r2=None, w2=0.0, out_f32=None, dst=None, dst_ch=0, c_pad=None):
        """one forward convolution: three launches into f32 scratch, then the epilogue kernel"""
        c = self.cv[name]
        sums, cs = self._three(src, c.hi, c.lo, src.B, src.H, src.W, cin, c.cout, c.n_pad, s)
        self._finish([t.data_ptr() for t in sums], src.npix, c.cout, s, sum_stride=cs, bias=c.bias.data_ptr(), act=act, s0=s0,
                     r1=r1.data_ptr() if r1 is not None else None, r1_stride=c.cout, w1=w1,
                     r2=r2.data_ptr() if r2 is not None else None, r2_stride=c.cout, w2=w2,
                     out_f32=out_f32.data_ptr() if out_f32 is not None else None, out32_stride=c.cout,
                     hi=dst.hi.ptr(dst_ch) if dst is not None else None, lo=dst.lo.ptr(dst_ch) if dst is not None else None,
                     out_stride=dst.C if dst is not None else 0, c_pad=c_pad)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, keep=False):
        """x: f32 NCHW cuda tensor [B, num_in_ch, h, w] -> f32 NCHW [B, num_out_ch, scale*h, scale*w].  keep: retain every activation
        pair for `backward`."""
        assert x.is_cuda and x.dtype == torch.float32
        if keep and not self.want_grad:
            raise RuntimeError("SplitBf16RRDBNet(want_grad=True) is needed to keep activations for backward")
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
        self._finish([x_nhwc.data_ptr()], npix, Cc, s, hi=in0.hi.ptr(), lo=in0.lo.ptr(), out_stride=self.cin_pad, c_pad=self.cin_pad)
        n_rdb = 3 * self.nb
        bufs = [_Pair(B, h, w, cw, dev) for _ in range(n_rdb if keep else 2)]
        buf_of = (lambda i: bufs[i]) if keep else (lambda i: bufs[i % 2])
        trunk = [torch.empty((npix, nf), dtype=torch.float32, device=dev) for _ in range(n_rdb + 1)]
        body_out = _Pair(B, h, w, nf, dev)
        self._conv("conv_first", in0, self.cin_pad, s, out_f32=trunk[0], dst=buf_of(0), dst_ch=0, c_pad=nf)
        for i in range(n_rdb):
            blk, j = divmod(i, 3)
            cur = buf_of(i)
            nxt = buf_of(i + 1) if i + 1 < n_rdb else body_out
            for k in range(1, 5):       # x_k = lrelu(conv_k(cat(x, x1..x_{k-1}))) -- rrdbnet_arch.py:39-42
                cin = nf + (k - 1) * g
                self._conv(f"body.{blk}.rdb{j + 1}.conv{k}", cur, cin, s, act=1, dst=cur, dst_ch=cin, c_pad=g)
            if j < 2:                   # x5 * 0.2 + x -- :43-44
                self._conv(f"body.{blk}.rdb{j + 1}.conv5", cur, cw, s, s0=0.2, r1=trunk[i], w1=1.0, out_f32=trunk[i + 1],
                           dst=nxt, dst_ch=0, c_pad=nf)
            else:                       # (x5 * 0.2 + x_rdb3) * 0.2 + x_rrdb -- :68
                self._conv(f"body.{blk}.rdb{j + 1}.conv5", cur, cw, s, s0=0.04, r1=trunk[i], w1=0.2, r2=trunk[3 * blk], w2=1.0,
                           out_f32=trunk[i + 1], dst=nxt, dst_ch=0, c_pad=nf)
        feat = _Pair(B, h, w, nf, dev)
        self._conv("conv_body", body_out, nf, s, r1=trunk[0], w1=1.0, dst=feat, c_pad=nf)      # feat = feat + body_feat -- :124-125
        src, hh, ww = feat, h, w
        up_in, up_out = [], []
        for u in range(self.n_up):     # conv_up(F.interpolate(feat, scale_factor=2, mode='nearest')) -- :127-134
            up = _Pair(B, hh * 2, ww * 2, nf, dev)
            for a, b_ in ((src.hi, up.hi), (src.lo, up.lo)):
                L.check(lb.ssr_upsample_nearest(a.ptr(), nf, b_.ptr(), nf, B, hh, ww, nf, 2, s))
            hh, ww = hh * 2, ww * 2
            out = _Pair(B, hh, ww, nf, dev)
            self._conv(f"conv_up{u + 1}", up, nf, s, act=1, dst=out, c_pad=nf)
            up_in.append(up)
            up_out.append(out)
            src = out
        hr = _Pair(B, hh, ww, nf, dev)
        self._conv("conv_hr", src, nf, s, act=1, dst=hr, c_pad=nf)                              # :136
        out = torch.empty((B * hh * ww, self.cout), dtype=torch.float32, device=dev)
        self._conv("conv_last", hr, nf, s, out_f32=out)
        if keep:
            self._saved = dict(B=B, h=h, w=w, H=hh, W=ww, in0=in0, bufs=bufs, body_out=body_out, up_in=up_in, up_out=up_out, hr=hr)
        return out.view(B, hh, ww, self.cout).permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ backward
    def _wgrad_bias(self, name, x, cx, dy, cy, B, H, W, s, x_ch=0, dy_ch=0):
        """dW += X^T dY (three bf16-pair launches into the conv's f32 accumulator), db += sum over pixels of dY"""
        lb = lib()
        for xb, db in ((x.hi, dy.hi), (x.lo, dy.hi), (x.hi, dy.lo)):
            a = self.wg.args(name, xb.ptr(x_ch), xb.stride, cx, db.ptr(dy_ch), db.stride, cy, B, H, W, 3, 1.0)
            L.check(lb.ssr_wgrad_tc(C.byref(a), s))
        gb = self.grads[f"{name}.bias"]
        for db in (dy.hi, dy.lo):
            L.check(lb.ssr_bias_grad(db.ptr(dy_ch), db.stride, B * H * W, cy, gb.data_ptr(), 1.0, s))

    def _dgrad(self, name, dy, B, H, W, s, dy_ch=0):
        """conv^T(dY): three launches over the mirrored operand -> ([3 f32 sum buffers], their pixel stride); channels = the conv's cin"""
        c = self.cv[name]
        return self._three(dy, c.dg_hi, c.dg_lo, B, H, W, c.cout_buf, c.cin, c.n_pad_dg, s, ch=dy_ch)

    @torch.no_grad()
    def backward(self, d_out):
        """d_out: f32 NCHW gradient of the forward output (of the last `forward(x, keep=True)`); returns {parameter name: f32 gradient}
        (the reference's autograd through rrdbnet_arch.py:116-137, evaluated in the split-bf16 form)."""
        sv = self._saved
        assert sv is not None, "call forward(x, keep=True) first"
        dev, nf, g, nb = self.device, self.nf, self.g, self.nb
        cw = nf + 4 * g
        B, h, w, H, W = sv["B"], sv["h"], sv["w"], sv["H"], sv["W"]
        s = cur_stream()
        lb = lib()
        for t in self.grads.values():
            t.zero_()
        self.wg.zero()
        f32 = lambda n, c: torch.empty((n, c), dtype=torch.float32, device=dev)
        P, p = B * H * W, B * h * w

        # ---- conv_last <- conv_hr
        d_nhwc = d_out.to(dev, torch.float32).permute(0, 2, 3, 1).contiguous()
        dy_last = _Pair(B, H, W, 16, dev)
        self._finish([d_nhwc.data_ptr()], P, self.cout, s, hi=dy_last.hi.ptr(), lo=dy_last.lo.ptr(), out_stride=16, c_pad=16)
        self._wgrad_bias("conv_last", sv["hr"], nf, dy_last, self.cout, B, H, W, s)
        sums, cs = self._dgrad("conv_last", dy_last, B, H, W, s)
        dy = _Pair(B, H, W, nf, dev)                                # dY of conv_hr = that, times LeakyReLU'(conv_hr output)
        self._finish([t.data_ptr() for t in sums], P, nf, s, sum_stride=cs, mask=sv["hr"].hi.ptr(), mask_stride=nf,
                     hi=dy.hi.ptr(), lo=dy.lo.ptr(), out_stride=nf)
        top = sv["up_out"][-1]
        self._wgrad_bias("conv_hr", top, nf, dy, nf, B, H, W, s)
        sums, cs = self._dgrad("conv_hr", dy, B, H, W, s)
        dy = _Pair(B, H, W, nf, dev)                                # dY of the last conv_up
        self._finish([t.data_ptr() for t in sums], P, nf, s, sum_stride=cs, mask=top.hi.ptr(), mask_stride=nf,
                     hi=dy.hi.ptr(), lo=dy.lo.ptr(), out_stride=nf)
        # ---- conv_up_n ... conv_up1, each behind a nearest x2 upsample
        hh, ww = H, W
        d_feat32 = None
        for u in range(self.n_up - 1, -1, -1):
            self._wgrad_bias(f"conv_up{u + 1}", sv["up_in"][u], nf, dy, nf, B, hh, ww, s)
            sums, cs = self._dgrad(f"conv_up{u + 1}", dy, B, hh, ww, s)
            full = f32(B * hh * ww, nf)
            self._finish([t.data_ptr() for t in sums], B * hh * ww, nf, s, sum_stride=cs, out_f32=full.data_ptr(), out32_stride=nf)
            hh, ww = hh // 2, ww // 2
            pooled = f32(B * hh * ww, nf)
            L.check(lb.ssr_sum_pool2x2_f32(full.data_ptr(), pooled.data_ptr(), B, hh, ww, nf, s))
            dy = _Pair(B, hh, ww, nf, dev)
            if u > 0:
                prev = sv["up_out"][u - 1]
                self._finish([pooled.data_ptr()], B * hh * ww, nf, s, mask=prev.hi.ptr(), mask_stride=nf,
                             hi=dy.hi.ptr(), lo=dy.lo.ptr(), out_stride=nf)
            else:
                self._finish([pooled.data_ptr()], p, nf, s, hi=dy.hi.ptr(), lo=dy.lo.ptr(), out_stride=nf)
                d_feat32 = pooled                                   # gradient w.r.t. feat = conv_first(x) + conv_body(body(.))
        d_feat = dy
        # ---- conv_body
        self._wgrad_bias("conv_body", sv["body_out"], nf, d_feat, nf, B, h, w, s)
        sums, cs = self._dgrad("conv_body", d_feat, B, h, w, s)
        g_rrdb = f32(p, nf)                                         # gradient w.r.t. the output of the current RRDB
        self._finish([t.data_ptr() for t in sums], p, nf, s, sum_stride=cs, out_f32=g_rrdb.data_ptr(), out32_stride=nf)
        # ---- the trunk, last block first
        G = torch.zeros((p, cw), dtype=torch.float32, device=dev)   # running gradient w.r.t. the block's dense buffer (x | x1..x4)
        g_out = f32(p, nf)
        dY5 = _Pair(B, h, w, nf, dev)
        dYk = _Pair(B, h, w, g, dev)
        for i in range(3 * nb - 1, -1, -1):
            blk, j = divmod(i, 3)
            cur = sv["bufs"][i]
            pre = f"body.{blk}.rdb{j + 1}"
            if j == 2:      # RRDB: out = rdb3_out * 0.2 + x  -> the third block's output gradient is 0.2 * g_rrdb
                self._finish([g_rrdb.data_ptr()], p, nf, s, s0=0.2, out_f32=g_out.data_ptr(), out32_stride=nf)
            # (j < 2: g_out already holds the gradient w.r.t. this block's output = the next block's input gradient)
            # block: out = x5 * 0.2 + x: G[0:64] = g_out, G[64:] = 0, dY5 = 0.2 * g_out
            G.zero_()
            self._finish([g_out.data_ptr()], p, nf, s, out_f32=G.data_ptr(), out32_stride=cw)
            self._finish([g_out.data_ptr()], p, nf, s, s0=0.2, hi=dY5.hi.ptr(), lo=dY5.lo.ptr(), out_stride=nf)
            self._wgrad_bias(f"{pre}.conv5", cur, cw, dY5, nf, B, h, w, s)
            sums, cs = self._dgrad(f"{pre}.conv5", dY5, B, h, w, s)
            self._finish([t.data_ptr() for t in sums], p, cw, s, sum_stride=cs, r1=G.data_ptr(), r1_stride=cw, w1=1.0,
                         out_f32=G.data_ptr(), out32_stride=cw)
            for k in range(4, 0, -1):
                nk = nf + (k - 1) * g                               # x_k lives in channels [nk, nk + g); conv_k reads [0, nk)
                # the slot is final: dY of conv_k = G[slot] * LeakyReLU'(x_k)
                self._finish([G.data_ptr() + 4 * nk], p, g, s, sum_stride=cw, mask=cur.hi.ptr(nk), mask_stride=cw,
                             hi=dYk.hi.ptr(), lo=dYk.lo.ptr(), out_stride=g)
                self._wgrad_bias(f"{pre}.conv{k}", cur, nk, dYk, g, B, h, w, s)
                sums, cs = self._dgrad(f"{pre}.conv{k}", dYk, B, h, w, s)
                self._finish([t.data_ptr() for t in sums], p, nk, s, sum_stride=cs, r1=G.data_ptr(), r1_stride=cw, w1=1.0,
                             out_f32=G.data_ptr(), out32_stride=cw)
            # gradient w.r.t. the block input = G[0:64]; the first block of an RRDB also receives the RRDB-level skip gradient
            if j == 0:
                self._finish([G.data_ptr()], p, nf, s, sum_stride=cw, r1=g_rrdb.data_ptr(), r1_stride=nf, w1=1.0,
                             out_f32=g_rrdb.data_ptr(), out32_stride=nf)          # = gradient w.r.t. the previous RRDB's output
            else:
                self._finish([G.data_ptr()], p, nf, s, sum_stride=cw, out_f32=g_out.data_ptr(), out32_stride=nf)
        # ---- conv_first: dY = trunk gradient + the long skip (feat = conv_first + conv_body(...))
        d_first = _Pair(B, h, w, nf, dev)
        self._finish([g_rrdb.data_ptr()], p, nf, s, r1=d_feat32.data_ptr(), r1_stride=nf, w1=1.0,
                     hi=d_first.hi.ptr(), lo=d_first.lo.ptr(), out_stride=nf)
        self._wgrad_bias("conv_first", sv["in0"], self.cin_pad, d_first, nf, B, h, w, s)
        self.wg.unpack(s)
        torch.cuda.current_stream().synchronize()
        return self.grads
