"""RRDBNet generator engine: the 351-conv forward (and backward) as a recorded plan of tcgen05 conv launches.

Dense-block layout: every ResidualDenseBlock owns ONE NHWC bf16 buffer of num_feat + 4*num_grow_ch channels;
conv_k reads channels [0, num_feat + (k-1)*grow) and writes its LeakyReLU'd output into the next `grow`
channels, so the four torch.cat of /root/reference/ssr/archs/rrdbnet_arch.py:39-42 never materialise.
conv5's epilogue applies x5*0.2 + x (:44) -- and for the third block of an RRDB also out*0.2 + x (:68) -- and
writes straight into channels [0, num_feat) of the NEXT block's buffer.
"""
import ctypes as C

import os

import torch

from . import _lib as L
from .ops import (Act, PackedConv, Packer, Plan, WgradSet, conv_args, cur_stream, lib, overlap_enabled, plan_wgrad, plan_wgrad_batch,
                  round_up)


class RRDBNetEngine:
    def __init__(self, params, num_in_ch, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32,
                 want_grad=True, grads=None, overlap=None):
        if scale not in (1, 2, 4, 8, 16):
            raise ValueError(f"RRDBNetEngine: scale {scale} (the reference builds 1, 2, 4, 8, 16: rrdbnet_arch.py:92-109)")
        if num_feat % 16 or num_grow_ch % 16:
            raise ValueError("num_feat and num_grow_ch must be multiples of 16")
        self.p = params
        self.device = next(iter(params.values())).device
        self.cin, self.cout, self.scale = num_in_ch, num_out_ch, scale
        self.nf, self.nb, self.g = num_feat, num_block, num_grow_ch
        # scale 2 / 1: pixel_unshuffle(x, 2 / 4) in front of conv_first (rrdbnet_arch.py:95-98, 117-120), fused into the ingest
        self.unshuffle = {1: 4, 2: 2}.get(scale, 1)
        self.cin_eff = num_in_ch * self.unshuffle ** 2
        if params["conv_first.weight"].shape[1] != self.cin_eff:
            raise ValueError(f"conv_first.weight has {params['conv_first.weight'].shape[1]} input channels, expected {self.cin_eff}")
        self.cin_pad = round_up(self.cin_eff, 16)
        self.want_grad = want_grad
        self.overlap = overlap_enabled("bwd", overlap)   # dense-block weight gradients on the side stream (_build_backward)
        self.n_up = {1: 2, 2: 2, 4: 2, 8: 3, 16: 4}[scale]
        nf, g = self.nf, self.g
        cv = {}

        def mk(name, cin_buf):
            cv[name] = PackedConv(params[f"{name}.weight"], params[f"{name}.bias"], cin_buf, want_grad, self.device)

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
        # The dense-block input-gradient chain keeps its running sum in tensor memory (ssr_conv_tc_chain_acc; SSR_DGRAD_TMEM=0 goes
        # back to the f32 sum in global memory).  Its layers are pure sums, so conv5's 0.2 (0.04 for the third block of an RRDB) is
        # folded into the packed input-gradient weights.
        self.dgrad_tmem = want_grad and os.environ.get("SSR_DGRAD_TMEM", "1") == "1"
        if self.dgrad_tmem:
            for i in range(num_block):
                for j in (1, 2, 3):
                    cv[f"body.{i}.rdb{j}.conv5"].dg_inv_scale = torch.full((1,), 25.0 if j == 3 else 5.0, dtype=torch.float32,
                                                                           device=self.device)
        self.packer = Packer(list(cv.values()), self.device)
        self._ws = {}
        # gradients: `grads` maps parameter name -> f32 tensor of the parameter's shape (views of a flat buffer)
        self.grads = grads
        self.wg = None
        if want_grad and grads is not None:
            self.wg = WgradSet(self.device)
            for name, c in cv.items():
                self.wg.add(name, c, c.cin_buf)
            self.wg.finalize(lambda name: grads[f"{name}.weight"])

    # ------------------------------------------------------------------ weights
    def repack(self, stream=None):
        self.packer.run(stream)

    # ------------------------------------------------------------------ workspaces
    def workspace(self, B, h, w, train):
        key = (B, h, w, bool(train))
        ws = self._ws.get(key)
        if ws is None:
            ws = _Workspace(self, B, h, w, train)
            self._ws[key] = ws
        return ws

    def forward(self, x, train=False, stream=None, ws=None):
        """x: f32 NCHW cuda tensor [B, num_in_ch, h, w] -> f32 NCHW [B, num_out_ch, scale*h, scale*w] (engine-owned)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B, Cc, h, w = x.shape
        assert Cc == self.cin
        f = self.unshuffle
        if h % f or w % f:
            raise ValueError(f"scale {self.scale}: input {h}x{w} is not divisible by the pixel_unshuffle factor {f}")
        if ws is None:
            ws = self.workspace(B, h // f, w // f, train)
        s = stream if stream is not None else cur_stream()
        if f == 1:
            L.check(lib().ssr_ingest_nchw(x.data_ptr(), L.SSR_F32, ws.in0.ptr(), ws.in0.stride, B, Cc, h, w, self.cin_pad,
                                          1.0, None, None, s))
        else:
            L.check(lib().ssr_ingest_nchw_unshuffle(x.data_ptr(), ws.in0.ptr(), ws.in0.stride, B, Cc, h, w, f, self.cin_pad, 1.0, s))
        ws.fwd.run(s)
        return ws.out

    def backward(self, d_out, B, h, w, stream=None, ws=None):
        """d_out: f32 NCHW gradient of the forward output; accumulates into self.grads (weights and biases).
        h, w: height / width of the forward INPUT."""
        assert self.wg is not None, "engine built without gradient buffers"
        if ws is None:
            ws = self.workspace(B, h // self.unshuffle, w // self.unshuffle, True)
        s = stream if stream is not None else cur_stream()
        if ws.bwd is None:
            ws.bwd = ws._build_backward(self)
        self.wg.zero()
        L.check(lib().ssr_ingest_nchw(d_out.data_ptr(), L.SSR_F32, ws.d_last.ptr(), 16, B, self.cout, ws.H, ws.W, 16, 1.0,
                                      None, None, s))
        ws.bwd.run(s)
        self.wg.unpack(s)


class _Workspace:
    """Buffers + the recorded forward plan for one (batch, height, width, train) shape."""

    def __init__(self, eng, B, h, w, train):
        dev = eng.device
        nf, g, nb = eng.nf, eng.g, eng.nb
        cw = nf + 4 * g
        self.B, self.h, self.w, self.train = B, h, w, train
        self.in0 = Act(B, h, w, eng.cin_pad, dev)
        n_rdb = 3 * nb
        if train:
            self.bufs = [Act(B, h, w, cw, dev) for _ in range(n_rdb)]
            rdb_buf = lambda i: self.bufs[i]
        else:
            # block 0 keeps its own buffer (conv_first's output is needed again after the trunk); the rest rotate
            # over 4 so an RRDB's input survives until its third block's epilogue has read it.
            self.bufs = [Act(B, h, w, cw, dev) for _ in range(min(n_rdb, 5))]
            rdb_buf = lambda i: self.bufs[0] if i == 0 else self.bufs[1 + (i - 1) % 4]
        self.rdb_buf = rdb_buf
        # f32 copy of the 64-channel trunk (the x of every x5*0.2 + x): keeps the 69 chained residual adds out of
        # bf16.  Forward-only, so it always rotates (slot 0 is block 0's, needed again by conv_body's skip add).
        self.trunk = [torch.empty((B, h, w, nf), dtype=torch.float32, device=dev) for _ in range(min(n_rdb, 5) + 1)]
        self.trunk_of = lambda i: self.trunk[0] if i == 0 else self.trunk[1 + (i - 1) % 4]
        self.body_out = Act(B, h, w, nf, dev)
        self.feat = Act(B, h, w, nf, dev)
        self.up_in, self.up_out = [], []
        hh, ww = h, w
        for _ in range(eng.n_up):
            hh, ww = hh * 2, ww * 2
            self.up_in.append(Act(B, hh, ww, nf, dev))
            self.up_out.append(Act(B, hh, ww, nf, dev))
        self.H, self.W = hh, ww
        self.hr = Act(B, hh, ww, nf, dev)
        self.out = torch.empty((B, eng.cout, hh, ww), dtype=torch.float32, device=dev)
        self.fwd = self._build_forward(eng)
        self.bwd = None

    def _build_forward(self, eng):
        B, h, w = self.B, self.h, self.w
        nf, g, nb = eng.nf, eng.g, eng.nb
        plan = Plan()
        bptr = lambda cvx: cvx.bias.data_ptr()

        c = eng.cv["conv_first"]
        b0 = self.rdb_buf(0)
        plan.conv(conv_args(self.in0.ptr(), B, h, w, self.in0.stride, eng.cin_pad, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                            bias=bptr(c), out=b0.ptr(0), out_stride=b0.stride,
                            out32=self.trunk_of(0).data_ptr(), out32_mode=L.OUT32_PLANAR4, out32_stride=nf))
        n_rdb = 3 * nb
        # Training batches (one cluster of 4 CTAs per image, <= one wave of the 148 SMs) run dense blocks as shared-memory-
        # resident launches: latency-bound, but launch + prologue + drain are paid once per launch -- and one launch takes up to
        # `fuse` consecutive blocks (ssr_rdb_resident_max_blocks: 12 = four RRDBs), the tile staying in shared memory across the block
        # boundaries.  Large inference batches are the opposite regime -- many waves of CTAs: there the persistent per-layer
        # kernel (tiles streamed back to back through double-buffered TMEM) keeps the tensor pipe busier than a chain that
        # serialises five layers per CTA.
        resident = self.train or B * max(1, (h * w) // 256) <= 2 * 148
        fuse = lib().ssr_rdb_resident_max_blocks(B, h, w) if resident else 1
        group, in_group = [], 0
        for i in range(n_rdb):
            blk, j = divmod(i, 3)
            cur = self.rdb_buf(i)
            nxt = self.rdb_buf(i + 1) if i + 1 < n_rdb else self.body_out
            t_cur, t_nxt = self.trunk_of(i), self.trunk_of(i + 1)
            block = []   # the five convs of the dense block run as one chained launch
            for k in range(1, 5):
                c = eng.cv[f"body.{blk}.rdb{j + 1}.conv{k}"]
                cin = nf + (k - 1) * g
                block.append(conv_args(cur.ptr(0), B, h, w, cur.stride, cin, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                       bias=bptr(c), act=1, out=cur.ptr(cin), out_stride=cur.stride))
            c = eng.cv[f"body.{blk}.rdb{j + 1}.conv5"]
            if j < 2:
                block.append(conv_args(cur.ptr(0), B, h, w, cur.stride, nf + 4 * g, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                    bias=bptr(c), s0=0.2, res1=t_cur.data_ptr(), res1_kind=L.SSR_F32_PLANAR4, res1_stride=nf, s1=1.0,
                                    out=nxt.ptr(0), out_stride=nxt.stride,
                                    out32=t_nxt.data_ptr(), out32_mode=L.OUT32_PLANAR4, out32_stride=nf))
            else:
                t_blk = self.trunk_of(3 * blk)
                # (x5*0.2 + x_rdb3)*0.2 + x_rrdb
                block.append(conv_args(cur.ptr(0), B, h, w, cur.stride, nf + 4 * g, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                    bias=bptr(c), s0=0.04, res1=t_cur.data_ptr(), res1_kind=L.SSR_F32_PLANAR4, res1_stride=nf, s1=0.2,
                                    res2=t_blk.data_ptr(), res2_kind=L.SSR_F32_PLANAR4, res2_stride=nf, s2=1.0,
                                    out=nxt.ptr(0), out_stride=nxt.stride,
                                    out32=t_nxt.data_ptr(), out32_mode=L.OUT32_PLANAR4, out32_stride=nf))
            if resident:
                group.extend(block)
                in_group += 1
                if in_group == fuse or i + 1 == n_rdb:
                    plan.chain(group)
                    group, in_group = [], 0
            else:
                for a in block:
                    plan.conv(a)
        c = eng.cv["conv_body"]
        plan.conv(conv_args(self.body_out.ptr(), B, h, w, nf, nf, c.packed.data_ptr(), 3, c.cout, c.n_pad, bias=bptr(c),
                            res1=self.trunk_of(0).data_ptr(), res1_kind=L.SSR_F32_PLANAR4, res1_stride=nf, s1=1.0,
                            out=self.feat.ptr(), out_stride=nf))
        src = self.feat
        hh, ww = h, w
        for u in range(eng.n_up):
            ui, uo = self.up_in[u], self.up_out[u]
            plan.add(lib().ssr_upsample_nearest, src.ptr(), src.stride, ui.ptr(), ui.stride, B, hh, ww, nf, 2)
            hh, ww = hh * 2, ww * 2
            c = eng.cv[f"conv_up{u + 1}"]
            plan.conv(conv_args(ui.ptr(), B, hh, ww, nf, nf, c.packed.data_ptr(), 3, c.cout, c.n_pad, bias=bptr(c), act=1,
                                out=uo.ptr(), out_stride=nf))
            src = uo
        c = eng.cv["conv_hr"]
        plan.conv(conv_args(src.ptr(), B, hh, ww, nf, nf, c.packed.data_ptr(), 3, c.cout, c.n_pad, bias=bptr(c), act=1,
                            out=self.hr.ptr(), out_stride=nf))
        c = eng.cv["conv_last"]
        plan.conv(conv_args(self.hr.ptr(), B, hh, ww, nf, nf, c.packed.data_ptr(), 3, c.cout, c.n_pad, bias=bptr(c),
                            out32=self.out.data_ptr(), out32_mode=L.OUT32_NCHW))
        return plan

    # ------------------------------------------------------------------ backward
    def _build_backward(self, eng):
        """Input-gradient (dgrad) chain + weight gradients for the whole generator.

        Per ResidualDenseBlock, with cur = the block's saved dense buffer (x | x1..x4):
          G32 (f32, 192 ch) accumulates the gradient w.r.t. every channel of the dense buffer; D (bf16) holds, per
          channel group, that gradient times LeakyReLU'(x_k) once the group is final = dY of conv_k.
          conv5^T first (scaled by the 0.2 of x5*0.2 + x), then conv4^T ... conv1^T, each adding into G32 through the
          epilogue's f32 residual and masking with the sign of the saved activation (rrdbnet_arch.py:37-44 reversed).
        """
        B, h, w, H, W = self.B, self.h, self.w, self.H, self.W
        nf, g, nb = eng.nf, eng.g, eng.nb
        cw = nf + 4 * g
        dev = eng.device
        wg = eng.wg
        F32 = L.SSR_F32_PLANAR4   # the running f32 gradients are epilogue-only buffers: quad-planar, coalesced per warp
        plan = Plan()
        grads = eng.grads
        self.d_last = Act(B, H, W, 16, dev)
        gA, gB = Act(B, H, W, nf, dev), Act(B, H, W, nf, dev)
        d_feat = Act(B, h, w, nf, dev)
        G32 = torch.empty((B, h, w, cw), dtype=torch.float32, device=dev)
        GO32 = torch.empty((B, h, w, nf), dtype=torch.float32, device=dev)
        # One resident launch takes the input-gradient chains of up to `fuse` consecutive blocks; their weight gradients follow the
        # launch, so everything those read must survive it: one dY buffer per block of a group, fuse + 1 rotating block-gradient
        # buffers (block i reads gR[(i+1) % m], writes gR[i % m]), and the RRDB-level gradient rotates over ceil(fuse / 3) + 1 buffers
        # (RRDB b reads gO[(b+1) % n], writes gO[b % n]).
        tmem_chain = bool(eng.dgrad_tmem and lib().ssr_conv_tc_chain_acc_supported(B, h, w, cw))
        fuse = lib().ssr_rdb_resident_max_blocks(B, h, w) if tmem_chain else 1
        # SSR_OVERLAP: a group's weight-gradient launches run on the side stream beside the NEXT group's input-gradient launch
        # (which occupies 128 of the 148 SMs), so what they read must survive that launch too: two sets of dY buffers
        # (group parity), 2 * fuse + 1 block-gradient and 2 * ceil(fuse / 3) + 2 RRDB-gradient buffers.  The launch after the
        # next waits for them (plan.join()).
        overlap = bool(tmem_chain and fuse > 1 and 3 * nb > fuse and eng.overlap)
        self.overlap_bwd = overlap
        n_sets = 2 if overlap else 1
        Dgs = [Act(B, h, w, cw, dev) for _ in range(n_sets * fuse)]
        m_r = n_sets * fuse + 1
        gR = [Act(B, h, w, nf, dev) for _ in range(m_r)]
        # a group spans up to ceil(fuse / 3) RRDB boundaries, and the deferred weight gradients read them all
        n_o = n_sets * ((fuse + 2) // 3) + n_sets + (1 if overlap else 0)
        gO = [Act(B, h, w, nf, dev) for _ in range(n_o)]
        d_first = Act(B, h, w, nf, dev)
        self._bwd_keep = [gA, gB, d_feat, G32, GO32, Dgs, gR, gO, d_first]

        def bias_grad(name, dy_ptr, dy_stride, npix, cy, scale=1.0):
            plan.add(lib().ssr_bias_grad, dy_ptr, dy_stride, npix, cy, grads[f"{name}.bias"].data_ptr(), scale)

        def wgrad(name, x_ptr, x_stride, cx, dy_ptr, dy_stride, cy, BB, HH, WW, scale=1.0, bias=True):
            plan_wgrad(plan, wg.args(name, x_ptr, x_stride, cx, dy_ptr, dy_stride, cy, BB, HH, WW, 3, scale))
            if bias:
                bias_grad(name, dy_ptr, dy_stride, BB * HH * WW, cy, scale)

        def bgrad_ptr(i, k):
            """bias gradient of conv k of dense block i (accumulated by the epilogue of the conv that produces its dY)"""
            blk, j = divmod(i, 3)
            return grads[f"body.{blk}.rdb{j + 1}.conv{k}.bias"].data_ptr()

        # ---- tail: conv_last <- conv_hr <- conv_up_n ... conv_up1
        c = eng.cv["conv_last"]
        plan.conv(conv_args(self.d_last.ptr(), B, H, W, 16, 16, c.packed_dg.data_ptr(), 3, nf, c.n_pad_dg,
                            mask=self.hr.ptr(), mask_stride=nf, mask_lo=0, out=gA.ptr(), out_stride=nf,
                            bias_grad=grads["conv_hr.bias"].data_ptr(), bias_grad_scale=1.0))   # gA = dY of conv_hr
        wgrad("conv_last", self.hr.ptr(), nf, nf, self.d_last.ptr(), 16, eng.cout, B, H, W)
        c = eng.cv["conv_hr"]
        top = self.up_out[-1]
        plan.conv(conv_args(gA.ptr(), B, H, W, nf, nf, c.packed_dg.data_ptr(), 3, nf, c.n_pad_dg,
                            mask=top.ptr(), mask_stride=nf, mask_lo=0, out=gB.ptr(), out_stride=nf,
                            bias_grad=grads[f"conv_up{eng.n_up}.bias"].data_ptr(), bias_grad_scale=1.0))   # gB = dY of the last conv_up
        wgrad("conv_hr", top.ptr(), nf, nf, gA.ptr(), nf, nf, B, H, W, bias=False)
        dy = gB            # dY of conv_up{n}
        hh, ww = H, W
        spare = gA
        for u in range(eng.n_up - 1, -1, -1):
            c = eng.cv[f"conv_up{u + 1}"]
            ui = self.up_in[u]
            d_ui = Act(B, hh, ww, nf, dev) if (spare.H, spare.W) != (hh, ww) else spare
            self._bwd_keep.append(d_ui)
            plan.conv(conv_args(dy.ptr(), B, hh, ww, nf, nf, c.packed_dg.data_ptr(), 3, nf, c.n_pad_dg,
                                out=d_ui.ptr(), out_stride=nf))
            wgrad(f"conv_up{u + 1}", ui.ptr(), nf, nf, dy.ptr(), nf, nf, B, hh, ww, bias=(u != eng.n_up - 1))
            hh, ww = hh // 2, ww // 2
            if u > 0:
                nxt = Act(B, hh, ww, nf, dev)
                self._bwd_keep.append(nxt)
                prev_act = self.up_out[u - 1]
                plan.add(lib().ssr_upsample_nearest_bwd, d_ui.ptr(), nf, nxt.ptr(), nf, B, hh, ww, nf, 2,
                         prev_act.ptr(), nf)
                dy = nxt
                spare = Act(B, hh, ww, nf, dev)
                self._bwd_keep.append(spare)
            else:
                plan.add(lib().ssr_upsample_nearest_bwd, d_ui.ptr(), nf, d_feat.ptr(), nf, B, hh, ww, nf, 2, None, 0)
        # ---- conv_body
        c = eng.cv["conv_body"]
        plan.conv(conv_args(d_feat.ptr(), B, h, w, nf, nf, c.packed_dg.data_ptr(), 3, nf, c.n_pad_dg,
                            out=gO[nb % n_o].ptr(), out_stride=nf, out32=GO32.data_ptr(), out32_mode=L.OUT32_PLANAR4, out32_stride=nf,
                            bias_grad=bgrad_ptr(3 * nb - 1, 5), bias_grad_scale=0.04))
        wgrad("conv_body", self.body_out.ptr(), nf, nf, d_feat.ptr(), nf, nf, B, h, w)
        # ---- the trunk, last block first
        pend_chain, pend_batches = [], []
        n_group = 0
        for i in range(3 * nb - 1, -1, -1):
            blk, j = divmod(i, 3)
            cur = self.bufs[i]
            pre = f"body.{blk}.rdb{j + 1}"
            c5 = eng.cv[f"{pre}.conv5"]
            # the block reads its incoming gradient from gR_in (written by block i+1) and hands its result on in gR_out: rotating,
            # because the weight-gradient launches (deferred to the end of the block / group) still read gR_in
            gR_in, gR_out = gR[(i + 1) % m_r], gR[i % m_r]
            gO_in, gO_b = gO[(blk + 1) % n_o], gO[blk % n_o]   # RRDB-level gradient: read by the third block, written by the first
            Dg = Dgs[(n_group % n_sets) * fuse + len(pend_batches)]
            if j == 2:
                xin, s0, r1, r1s, s1 = gO_in, 0.04, GO32.data_ptr(), nf, 0.2
            else:
                xin, s0, r1, r1s, s1 = gR_in, 0.2, G32.data_ptr(), cw, 1.0
            if tmem_chain:
                # running sum in tensor memory: layer k emits only its finished dY slot (masked bf16 + bias gradient); the last
                # layer emits the block-input gradient = sum + incoming gradient(s).  G32 shrinks to the 64 trunk channels.
                s0w = 0.04 if j == 2 else 0.2
                achain = [conv_args(xin.ptr(), B, h, w, nf, nf, c5.packed_dg.data_ptr(), 3, cw, c5.n_pad_dg,
                                    mask=cur.ptr(), mask_stride=cw, mask_lo=cw - g, out_lo=cw - g, bias_grad=bgrad_ptr(i, 4),
                                    out=Dg.ptr(), out_stride=cw)]
                batch = [wg.args(f"{pre}.conv5", cur.ptr(), cw, cw, xin.ptr(), nf, nf, B, h, w, 3, s0w)]
                for k in range(4, 0, -1):
                    ck = eng.cv[f"{pre}.conv{k}"]
                    nk = nf + (k - 1) * g
                    dyk = Dg.ptr(nk)
                    if k > 1:
                        achain.append(conv_args(dyk, B, h, w, cw, g, ck.packed_dg.data_ptr(), 3, nk, ck.n_pad_dg,
                                                mask=cur.ptr(), mask_stride=cw, mask_lo=nk - g, out_lo=nk - g,
                                                bias_grad=bgrad_ptr(i, k - 1), out=Dg.ptr(), out_stride=cw))
                    else:
                        res = dict(res1=GO32.data_ptr(), res1_kind=F32, res1_stride=nf, s1=0.2) if j == 2 else \
                            dict(res1=G32.data_ptr(), res1_kind=F32, res1_stride=cw, s1=1.0)
                        if j == 0:
                            res.update(res2=GO32.data_ptr(), res2_kind=F32, res2_stride=nf, s2=1.0)
                        dst, dst32 = (gO_b, GO32) if j == 0 else (gR_out, G32)
                        achain.append(conv_args(dyk, B, h, w, cw, g, ck.packed_dg.data_ptr(), 3, nk, ck.n_pad_dg,
                                                out=dst.ptr(), out_stride=nf, out32=dst32.data_ptr(), out32_mode=L.OUT32_PLANAR4,
                                                out32_stride=nf, bias_grad=bgrad_ptr(i - 1, 5) if i > 0 else None,
                                                bias_grad_scale=0.04 if j == 0 else 0.2, **res))
                    batch.append(wg.args(f"{pre}.conv{k}", cur.ptr(), cw, nk, dyk, cw, g, B, h, w, 3, 1.0))
                pend_chain.extend(achain)
                pend_batches.append(batch)
                if len(pend_batches) == fuse or i == 0:
                    plan.chain_acc(pend_chain)       # ONE launch: the input-gradient chains of the group's blocks
                    if overlap:
                        if n_group:
                            plan.join()              # the NEXT launch rewrites the dY set group n_group - 1's weight gradients read
                        plan.fork()
                        with plan.side():
                            for bt in pend_batches:
                                plan_wgrad_batch(plan, bt)
                    else:
                        for bt in pend_batches:      # then each block's five weight gradients in one launch
                            plan_wgrad_batch(plan, bt)
                    pend_chain, pend_batches = [], []
                    n_group += 1
                continue
            if eng.dgrad_tmem:
                s0 = 1.0   # the factor lives in the packed weights (see RRDBNetEngine.__init__); the wgrad scale below keeps it
            # The five input-gradient convs of the block: one chained launch (each reads the dY slot the previous one wrote).
            # Every layer adds into ALL its f32 channels (G32) but only its top g-channel slot is read again as bf16 -- it is
            # the finished dY of the conv below, so the masked bf16 store is limited to that slot (out_lo) and its pixel
            # sum is accumulated as that conv's bias gradient on the way out.
            dchain = [conv_args(xin.ptr(), B, h, w, nf, nf, c5.packed_dg.data_ptr(), 3, cw, c5.n_pad_dg, s0=s0,
                                res1=r1, res1_kind=F32, res1_stride=r1s, s1=s1, res1_cmax=nf,
                                mask=cur.ptr(), mask_stride=cw, mask_lo=cw - g, out_lo=cw - g, bias_grad=bgrad_ptr(i, 4),
                                out=Dg.ptr(), out_stride=cw, out32=G32.data_ptr(),
                                out32_mode=L.OUT32_PLANAR4_ACC if j < 2 else L.OUT32_PLANAR4, out32_stride=cw)]
            batch = [wg.args(f"{pre}.conv5", cur.ptr(), cw, cw, xin.ptr(), nf, nf, B, h, w, 3, 0.04 if j == 2 else 0.2)]
            for k in range(4, 0, -1):
                ck = eng.cv[f"{pre}.conv{k}"]
                nk = nf + (k - 1) * g
                dyk = Dg.ptr(nk)
                if k > 1:
                    dchain.append(conv_args(dyk, B, h, w, cw, g, ck.packed_dg.data_ptr(), 3, nk, ck.n_pad_dg,
                                        res1=G32.data_ptr(), res1_kind=F32, res1_stride=cw, s1=1.0,
                                        mask=cur.ptr(), mask_stride=cw, mask_lo=nk - g, out_lo=nk - g,
                                        bias_grad=bgrad_ptr(i, k - 1),
                                        out=Dg.ptr(), out_stride=cw, out32=G32.data_ptr(), out32_mode=L.OUT32_PLANAR4_ACC,
                                        out32_stride=cw))
                elif j > 0:
                    dchain.append(conv_args(dyk, B, h, w, cw, g, ck.packed_dg.data_ptr(), 3, nk, ck.n_pad_dg,
                                        res1=G32.data_ptr(), res1_kind=F32, res1_stride=cw, s1=1.0,
                                        out=gR_out.ptr(), out_stride=nf, out32=G32.data_ptr(), out32_mode=L.OUT32_PLANAR4,
                                        out32_stride=cw, bias_grad=bgrad_ptr(i - 1, 5), bias_grad_scale=0.2))
                else:
                    # first block of the RRDB: add the RRDB-level skip gradient and hand over to the previous RRDB
                    dchain.append(conv_args(dyk, B, h, w, cw, g, ck.packed_dg.data_ptr(), 3, nk, ck.n_pad_dg,
                                        res1=G32.data_ptr(), res1_kind=F32, res1_stride=cw, s1=1.0,
                                        res2=GO32.data_ptr(), res2_kind=F32, res2_stride=nf, s2=1.0,
                                        out=gO_b.ptr(), out_stride=nf, out32=GO32.data_ptr(), out32_mode=L.OUT32_PLANAR4,
                                        out32_stride=nf, bias_grad=bgrad_ptr(i - 1, 5) if i > 0 else None,
                                        bias_grad_scale=0.04))
                batch.append(wg.args(f"{pre}.conv{k}", cur.ptr(), cw, nk, dyk, cw, g, B, h, w, 3, 1.0))
            plan.chain(dchain)
            # all five weight gradients of the block in ONE launch (they only need the block's finished dY slots)
            plan_wgrad_batch(plan, batch)
        # ---- conv_first: dY = trunk gradient + the long skip (feat = conv_first + conv_body(...))
        plan.add(lib().ssr_axpby, gO[0].ptr(), nf, 1.0, d_feat.ptr(), nf, 1.0, None, 0, 0, d_first.ptr(), nf, B * h * w, nf)
        wgrad("conv_first", self.in0.ptr(), self.in0.stride, eng.cin_pad, d_first.ptr(), nf, nf, B, h, w)
        return plan
