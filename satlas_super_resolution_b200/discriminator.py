"""U-Net spectral-norm discriminator engine (forward, input gradient, weight gradients) over the C ABI.

Mirrors /root/reference/ssr/archs/discriminator_arch.py:42-71:
  conv0 (3x3, bias) -> conv1..3 (4x4 stride 2, spectral norm; implicit GEMM: the TMA box gathers every second pixel per tap,
  no im2col columns; input gradient = four 2x2 parity-class convs over dY scattered to (2y+oy, 2x+ox)) -> bilinear x2 -> conv4 (+x2)
  -> bilinear x2 -> conv5 (+x1) -> bilinear x2 -> conv6 (+x0) -> conv7, conv8 -> conv9 (bias) -> logits.
The three `x = x + skip` adds are folded into the consumer (the bilinear kernel's second source / one axpby), so the
LeakyReLU outputs stay available unmodified as the masks of the backward pass.
Spectral norm: one batched power iteration per training-mode forward (even with frozen weights), sigma written to
device memory and consumed by the weight packer as 1/sigma -- torch.nn.utils.spectral_norm semantics.
"""
import ctypes as C

import torch

from . import _lib as L
from ._protos import SnDesc
from .ops import (Act, PackedConv, Packer, Plan, WgradSet, conv_args, cur_stream, dgrad_s2_class_ptr, lib, plan_wgrad, round_up)

SN_LAYERS = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv8"]


class UNetDiscEngine:
    def __init__(self, params, num_in_ch, num_feat=64, skip_connection=True, grads=None):
        """params: name -> cuda f32 tensor with the reference state_dict keys (weight_orig / weight_u / weight_v for the
        normalised convs).  grads: name -> f32 tensor for conv0/9 weight+bias and conv1..8 weight_orig (or None)."""
        self.p, self.grads = params, grads
        self.device = params["conv0.weight"].device
        self.cin, self.nf, self.skip = num_in_ch, num_feat, skip_connection
        self.cin_pad = round_up(num_in_ch, 16)
        nf, dev = num_feat, self.device
        want = grads is not None
        self.sigma = torch.ones(8, dtype=torch.float32, device=dev)
        sg = lambda i: self.sigma[i:i + 1]
        cv = {}
        cv["conv0"] = PackedConv(params["conv0.weight"], params["conv0.bias"], self.cin_pad, True, dev)
        for i, name in enumerate(("conv1", "conv2", "conv3")):
            w = params[f"{name}.weight_orig"]
            cv[name] = PackedConv(w, None, w.shape[1], True, dev, inv_scale=sg(i))
        for i, name in enumerate(("conv4", "conv5", "conv6", "conv7", "conv8")):
            w = params[f"{name}.weight_orig"]
            cv[name] = PackedConv(w, None, w.shape[1], True, dev, inv_scale=sg(3 + i))
        cv["conv9"] = PackedConv(params["conv9.weight"], params["conv9.bias"], nf, True, dev)
        self.cv = cv
        self.packer = Packer(list(cv.values()), dev)
        # spectral-norm table
        self.geff = {}
        descs = []
        self._sn_scratch = []
        for i, name in enumerate(SN_LAYERS):
            w = params[f"{name}.weight_orig"]
            rows, cols = w.shape[0], w[0].numel()
            scratch = torch.zeros(cols + rows + 4, dtype=torch.float32, device=dev)
            self._sn_scratch.append(scratch)
            d = SnDesc()
            d.w, d.u, d.v = w.data_ptr(), params[f"{name}.weight_u"].data_ptr(), params[f"{name}.weight_v"].data_ptr()
            d.sigma, d.scratch = self.sigma.data_ptr() + 4 * i, scratch.data_ptr()
            if want:
                self.geff[name] = torch.zeros_like(w)
                d.geff, d.grad = self.geff[name].data_ptr(), grads[f"{name}.weight_orig"].data_ptr()
            d.rows, d.cols = rows, cols
            descs.append(d)
        arr = (SnDesc * 8)(*descs)
        self.sn_table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.wg = None
        if want:
            self.wg = WgradSet(dev)
            self.wg.add("conv0", cv["conv0"], self.cin_pad)
            for name in ("conv1", "conv2", "conv3"):
                self.wg.add(name, cv[name], cv[name].cin)
            for name in ("conv4", "conv5", "conv6", "conv7", "conv8", "conv9"):
                self.wg.add(name, cv[name], cv[name].cin)

            def grad_of(name):
                if name in SN_LAYERS:
                    return self.geff[name]
                return grads[f"{name}.weight"]
            # the normalised convs unpack into geff (overwrite); conv0 / conv9 accumulate straight into their grads
            self.wg.finalize(grad_of, accumulate=1)
            self._fix_unpack_modes()
        self._ws = {}

    def _fix_unpack_modes(self):
        """geff must be overwritten (it is per backward pass), plain grads accumulated."""
        from ._protos import UnpackDesc
        raw = bytearray(self.wg.table.cpu().numpy().tobytes())
        arr = (UnpackDesc * self.wg.n).from_buffer(raw)
        for idx, (name, *_rest) in enumerate(self.wg.items):
            arr[idx].accumulate = 0 if name in SN_LAYERS else 1
        self.wg.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)

    # ------------------------------------------------------------------
    def prepare_weights(self, training, stream=None):
        """spectral-norm sigma (with / without power iteration) then repack every conv as W / sigma."""
        s = stream if stream is not None else cur_stream()
        L.check(lib().ssr_spectral_norm(self.sn_table.data_ptr(), 8, 1 if training else 0, 1e-12, s))
        self.packer.run(s)

    def workspace(self, B, H, W):
        key = (B, H, W)
        ws = self._ws.get(key)
        if ws is None:
            ws = _DWorkspace(self, B, H, W)
            self._ws[key] = ws
        return ws

    def forward(self, ws, training=True, stream=None, prepared=False):
        """ws.x_in must hold the NHWC bf16 input; returns the engine-owned f32 [B,1,H,W] logits.
        prepared: prepare_weights(training) has already been issued for this pass (and is ordered before `stream`)."""
        s = stream if stream is not None else cur_stream()
        if not prepared:
            self.prepare_weights(training, s)
        ws.fwd.run(s)
        return ws.logits

    def backward(self, ws, d_logits, need_wgrad=True, need_dinput=False, stream=None):
        """d_logits: f32 [B,1,H,W].  Accumulates parameter gradients; the input gradient lands in ws.d_in (NHWC bf16)."""
        s = stream if stream is not None else cur_stream()
        if ws.bwd is None:
            ws.build_backward(self)
        B, H, W = ws.B, ws.H, ws.W
        L.check(lib().ssr_ingest_nchw(d_logits.data_ptr(), L.SSR_F32, ws.d9.ptr(), 16, B, 1, H, W, 16, 1.0, None, None, s))
        if need_wgrad:
            self.wg.zero()
            ws.bwd_full.run(s)
            self.wg.unpack(s)
            L.check(lib().ssr_spectral_norm_bwd(self.sn_table.data_ptr(), 8, s))
        else:
            ws.bwd.run(s)
        if need_dinput:
            ws.bwd_input.run(s)


class _DWorkspace:
    def __init__(self, eng, B, H, W):
        assert H % 8 == 0 and W % 8 == 0, "discriminator input must be divisible by 8"
        dev, nf = eng.device, eng.nf
        self.B, self.H, self.W = B, H, W
        A = lambda hh, ww, c: Act(B, hh, ww, c, dev)
        self.x_in = Act(B, H, W, eng.cin_pad, dev, zero=True)
        self.x0 = A(H, W, nf)
        self.x1 = A(H // 2, W // 2, nf * 2)
        self.x2 = A(H // 4, W // 4, nf * 4)
        self.x3 = A(H // 8, W // 8, nf * 8)
        self.x3u = A(H // 4, W // 4, nf * 8)
        self.a4 = A(H // 4, W // 4, nf * 4)
        self.x4u = A(H // 2, W // 2, nf * 4)
        self.a5 = A(H // 2, W // 2, nf * 2)
        self.x5u = A(H, W, nf * 2)
        self.a6 = A(H, W, nf)
        self.x6 = A(H, W, nf)
        self.a7 = A(H, W, nf)
        self.a8 = A(H, W, nf)
        self.logits = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        self.fwd = self._build_forward(eng)
        self.bwd = None

    def _build_forward(self, eng):
        B, H, W, nf = self.B, self.H, self.W, eng.nf
        plan = Plan()
        cv = eng.cv
        lb = lib()
        c = cv["conv0"]
        plan.conv(conv_args(self.x_in.ptr(), B, H, W, self.x_in.stride, eng.cin_pad, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                            bias=c.bias.data_ptr(), act=1, out=self.x0.ptr(), out_stride=nf))

        def strided(name, src, dst, hh, ww, cin):
            """4 x 4 stride-2 pad-1 conv + LeakyReLU (discriminator_arch.py:45-47) straight from the NHWC input: (hh, ww) = input size"""
            c = cv[name]
            plan.conv(conv_args(src.ptr(), B, hh, ww, src.stride, cin, c.packed.data_ptr(), 4, c.cout, c.n_pad, act=1,
                                out=dst.ptr(), out_stride=dst.stride, stride=2))

        strided("conv1", self.x0, self.x1, H, W, nf)
        strided("conv2", self.x1, self.x2, H // 2, W // 2, nf * 2)
        strided("conv3", self.x2, self.x3, H // 4, W // 4, nf * 4)

        def conv3(name, src, dst, hh, ww):
            c = cv[name]
            plan.conv(conv_args(src.ptr(), B, hh, ww, src.stride, c.cin, c.packed.data_ptr(), 3, c.cout, c.n_pad, act=1,
                                out=dst.ptr(), out_stride=dst.stride))

        sk = eng.skip
        plan.add(lb.ssr_upsample_bilinear2x, self.x3.ptr(), self.x3.stride, None, 0, self.x3u.ptr(), self.x3u.stride, B,
                 H // 8, W // 8, nf * 8)
        conv3("conv4", self.x3u, self.a4, H // 4, W // 4)
        plan.add(lb.ssr_upsample_bilinear2x, self.a4.ptr(), self.a4.stride, self.x2.ptr() if sk else None, self.x2.stride,
                 self.x4u.ptr(), self.x4u.stride, B, H // 4, W // 4, nf * 4)
        conv3("conv5", self.x4u, self.a5, H // 2, W // 2)
        plan.add(lb.ssr_upsample_bilinear2x, self.a5.ptr(), self.a5.stride, self.x1.ptr() if sk else None, self.x1.stride,
                 self.x5u.ptr(), self.x5u.stride, B, H // 2, W // 2, nf * 2)
        conv3("conv6", self.x5u, self.a6, H, W)
        plan.add(lb.ssr_axpby, self.a6.ptr(), nf, 1.0, self.x0.ptr() if sk else None, nf, 1.0, None, 0, 0, self.x6.ptr(), nf,
                 B * H * W, nf)
        conv3("conv7", self.x6, self.a7, H, W)
        conv3("conv8", self.a7, self.a8, H, W)
        c = cv["conv9"]
        plan.conv(conv_args(self.a8.ptr(), B, H, W, nf, nf, c.packed.data_ptr(), 3, 1, c.n_pad, bias=c.bias.data_ptr(),
                            out32=self.logits.data_ptr(), out32_mode=L.OUT32_NCHW))
        return plan

    def build_backward(self, eng):
        """Three plans: dgrad-only chain (frozen D in the generator step), dgrad+wgrad chain, and the final conv0^T that
        produces the gradient w.r.t. the discriminator input."""
        B, H, W, nf = self.B, self.H, self.W, eng.nf
        dev = eng.device
        cv, lb, wg = eng.cv, lib(), eng.wg
        A = lambda hh, ww, c: Act(B, hh, ww, c, dev)
        self.d9 = A(H, W, 16)
        g8, g7, gx6, g6 = A(H, W, nf), A(H, W, nf), A(H, W, nf), A(H, W, nf)
        gx5u = A(H, W, nf * 2)
        gs5, g5 = A(H // 2, W // 2, nf * 2), A(H // 2, W // 2, nf * 2)
        gx4u = A(H // 2, W // 2, nf * 4)
        gs4, g4 = A(H // 4, W // 4, nf * 4), A(H // 4, W // 4, nf * 4)
        gx3u = A(H // 4, W // 4, nf * 8)
        gx3, g3 = A(H // 8, W // 8, nf * 8), A(H // 8, W // 8, nf * 8)
        g2 = A(H // 4, W // 4, nf * 4)
        g1 = A(H // 2, W // 2, nf * 2)
        g0 = A(H, W, nf)
        self.d_in = A(H, W, eng.cin_pad)
        self._keep = [g8, g7, gx6, g6, gx5u, gs5, g5, gx4u, gs4, g4, gx3u, gx3, g3, g2, g1, g0]
        sk = eng.skip
        grads = eng.grads

        def build(with_wgrad):
            plan = Plan()

            def wgrad(name, x_ptr, x_stride, cx, dy_ptr, dy_stride, cy, BB, hh, ww, r=3):
                if not with_wgrad:
                    return
                from .ops import WgradSet  # noqa: F401
                a = wg.args(name, x_ptr, x_stride, cx, dy_ptr, dy_stride, cy, BB, hh, ww, r)
                plan_wgrad(plan, a)

            def dgrad3(name, x, cin_x, dst, hh, ww, mask=None):
                c = cv[name]
                plan.conv(conv_args(x.ptr(), B, hh, ww, x.stride, cin_x, c.packed_dg.data_ptr(), 3, c.cin, c.n_pad_dg,
                                    mask=mask.ptr() if mask is not None else None,
                                    mask_stride=mask.stride if mask is not None else 0, mask_lo=0,
                                    out=dst.ptr(), out_stride=dst.stride))

            # conv9 .. conv7
            dgrad3("conv9", self.d9, 16, g8, H, W, mask=self.a8)
            wgrad("conv9", self.a8.ptr(), nf, nf, self.d9.ptr(), 16, 1, B, H, W)
            if with_wgrad:
                plan.add(lb.ssr_bias_grad, self.d9.ptr(), 16, B * H * W, 1, grads["conv9.bias"].data_ptr(), 1.0)
            dgrad3("conv8", g8, nf, g7, H, W, mask=self.a7)
            wgrad("conv8", self.a7.ptr(), nf, nf, g8.ptr(), nf, nf, B, H, W)
            dgrad3("conv7", g7, nf, gx6, H, W)
            wgrad("conv7", self.x6.ptr(), nf, nf, g7.ptr(), nf, nf, B, H, W)
            # x6 = a6 + x0
            plan.add(lb.ssr_axpby, gx6.ptr(), nf, 1.0, None, 0, 0.0, self.a6.ptr(), nf, 0, g6.ptr(), nf, B * H * W, nf)
            dgrad3("conv6", g6, nf, gx5u, H, W)
            wgrad("conv6", self.x5u.ptr(), nf * 2, nf * 2, g6.ptr(), nf, nf, B, H, W)
            plan.add(lb.ssr_upsample_bilinear2x_bwd, gx5u.ptr(), nf * 2, gs5.ptr(), nf * 2, B, H // 2, W // 2, nf * 2)
            plan.add(lb.ssr_axpby, gs5.ptr(), nf * 2, 1.0, None, 0, 0.0, self.a5.ptr(), nf * 2, 0, g5.ptr(), nf * 2,
                     B * (H // 2) * (W // 2), nf * 2)
            dgrad3("conv5", g5, nf * 2, gx4u, H // 2, W // 2)
            wgrad("conv5", self.x4u.ptr(), nf * 4, nf * 4, g5.ptr(), nf * 2, nf * 2, B, H // 2, W // 2)
            plan.add(lb.ssr_upsample_bilinear2x_bwd, gx4u.ptr(), nf * 4, gs4.ptr(), nf * 4, B, H // 4, W // 4, nf * 4)
            plan.add(lb.ssr_axpby, gs4.ptr(), nf * 4, 1.0, None, 0, 0.0, self.a4.ptr(), nf * 4, 0, g4.ptr(), nf * 4,
                     B * (H // 4) * (W // 4), nf * 4)
            dgrad3("conv4", g4, nf * 4, gx3u, H // 4, W // 4)
            wgrad("conv4", self.x3u.ptr(), nf * 8, nf * 8, g4.ptr(), nf * 4, nf * 4, B, H // 4, W // 4)
            plan.add(lb.ssr_upsample_bilinear2x_bwd, gx3u.ptr(), nf * 8, gx3.ptr(), nf * 8, B, H // 8, W // 8, nf * 8)
            plan.add(lb.ssr_axpby, gx3.ptr(), nf * 8, 1.0, None, 0, 0.0, self.x3.ptr(), nf * 8, 0, g3.ptr(), nf * 8,
                     B * (H // 8) * (W // 8), nf * 8)

            def strided_bwd(name, gy, cy, src, dst, hh, ww, cin, skip_grad, act_in, bias_grad=None):
                """gy: dY of the strided conv [B, hh/2, ww/2, cy]; (hh, ww) = INPUT size of the conv; src = its input activation;
                dst = dY of the producer = (conv^T(gy) + skip gradient) * LeakyReLU'(act_in).  The transposed conv runs as four
                2 x 2 convs over gy, one per parity (oy, ox) of the output pixel: dx[2u+oy, 2v+ox] reads gy rows u - (1-oy) + {0, 1}."""
                c = cv[name]
                res = {}
                if bias_grad is not None:     # dst = dY of the producing conv: its bias gradient is the pixel sum, taken in the epilogues
                    res = dict(bias_grad=bias_grad.data_ptr(), bias_grad_scale=1.0)
                if skip_grad is not None and sk:
                    res.update(res1=skip_grad.ptr(), res1_kind=L.SSR_BF16, res1_stride=skip_grad.stride, s1=1.0)
                for cls in range(4):
                    oy, ox = divmod(cls, 2)
                    plan.conv(conv_args(gy.ptr(), B, hh // 2, ww // 2, gy.stride, cy, dgrad_s2_class_ptr(c, cls), 2, cin, c.n_pad_dg,
                                        pad_y=1 - oy, pad_x=1 - ox, out_oy=oy, out_ox=ox,
                                        mask=act_in.ptr(), mask_stride=act_in.stride, mask_lo=0,
                                        out=dst.ptr(), out_stride=dst.stride, **res))
                wgrad(name, src.ptr(), src.stride, cin, gy.ptr(), gy.stride, cy, B, hh, ww, r=4)

            strided_bwd("conv3", g3, nf * 8, self.x2, g2, H // 4, W // 4, nf * 4, gs4, self.x2)
            strided_bwd("conv2", g2, nf * 4, self.x1, g1, H // 2, W // 2, nf * 2, gs5, self.x1)
            strided_bwd("conv1", g1, nf * 2, self.x0, g0, H, W, nf, gx6, self.x0,
                        bias_grad=grads["conv0.bias"] if with_wgrad else None)
            wgrad("conv0", self.x_in.ptr(), self.x_in.stride, eng.cin_pad, g0.ptr(), nf, nf, B, H, W)
            return plan

        self.bwd = build(False)
        self.bwd_full = build(True) if wg is not None else None
        c = cv["conv0"]
        self.bwd_input = Plan()
        self.bwd_input.conv(conv_args(g0.ptr(), B, H, W, nf, nf, c.packed_dg.data_ptr(), 3, eng.cin, c.n_pad_dg,
                                      out=self.d_in.ptr(), out_stride=self.d_in.stride))
