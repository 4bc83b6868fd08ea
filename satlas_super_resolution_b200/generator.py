"""RRDBNet generator engine: the 351-conv forward (and backward) as a recorded plan of tcgen05 conv launches.

Dense-block layout: every ResidualDenseBlock owns ONE NHWC bf16 buffer of num_feat + 4*num_grow_ch channels;
conv_k reads channels [0, num_feat + (k-1)*grow) and writes its LeakyReLU'd output into the next `grow`
channels, so the four torch.cat of /root/reference/ssr/archs/rrdbnet_arch.py:39-42 never materialise.
conv5's epilogue applies x5*0.2 + x (:44) -- and for the third block of an RRDB also out*0.2 + x (:68) -- and
writes straight into channels [0, num_feat) of the NEXT block's buffer.
"""
import ctypes as C

import torch

from . import _lib as L
from .ops import Act, PackedConv, Packer, Plan, conv_args, cur_stream, lib, round_up


class RRDBNetEngine:
    def __init__(self, params, num_in_ch, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32,
                 want_grad=True):
        if scale not in (4, 8, 16):
            raise NotImplementedError("RRDBNetEngine: scale 1/2 (pixel_unshuffle front-end) is not built yet")
        if num_feat % 16 or num_grow_ch % 16:
            raise ValueError("num_feat and num_grow_ch must be multiples of 16")
        self.p = params
        self.device = next(iter(params.values())).device
        self.cin, self.cout, self.scale = num_in_ch, num_out_ch, scale
        self.nf, self.nb, self.g = num_feat, num_block, num_grow_ch
        self.cin_pad = round_up(num_in_ch, 16)
        self.want_grad = want_grad
        self.n_up = {4: 2, 8: 3, 16: 4}[scale]
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
        self.packer = Packer(list(cv.values()), self.device)
        self._ws = {}

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

    def forward(self, x, train=False, stream=None):
        """x: f32 NCHW cuda tensor [B, num_in_ch, h, w] -> f32 NCHW [B, num_out_ch, scale*h, scale*w] (engine-owned)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B, Cc, h, w = x.shape
        assert Cc == self.cin
        ws = self.workspace(B, h, w, train)
        s = stream if stream is not None else cur_stream()
        L.check(lib().ssr_ingest_nchw(x.data_ptr(), L.SSR_F32, ws.in0.ptr(), ws.in0.stride, B, Cc, h, w, self.cin_pad,
                                      1.0, None, None, s))
        ws.fwd.run(s)
        return ws.out


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

    def _build_forward(self, eng):
        B, h, w = self.B, self.h, self.w
        nf, g, nb = eng.nf, eng.g, eng.nb
        plan = Plan()
        bptr = lambda cvx: cvx.bias.data_ptr()

        c = eng.cv["conv_first"]
        b0 = self.rdb_buf(0)
        plan.conv(conv_args(self.in0.ptr(), B, h, w, self.in0.stride, eng.cin_pad, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                            bias=bptr(c), out=b0.ptr(0), out_stride=b0.stride,
                            out32=self.trunk_of(0).data_ptr(), out32_mode=L.OUT32_NHWC, out32_stride=nf))
        n_rdb = 3 * nb
        for i in range(n_rdb):
            blk, j = divmod(i, 3)
            cur = self.rdb_buf(i)
            nxt = self.rdb_buf(i + 1) if i + 1 < n_rdb else self.body_out
            t_cur, t_nxt = self.trunk_of(i), self.trunk_of(i + 1)
            for k in range(1, 5):
                c = eng.cv[f"body.{blk}.rdb{j + 1}.conv{k}"]
                cin = nf + (k - 1) * g
                plan.conv(conv_args(cur.ptr(0), B, h, w, cur.stride, cin, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                    bias=bptr(c), act=1, out=cur.ptr(cin), out_stride=cur.stride))
            c = eng.cv[f"body.{blk}.rdb{j + 1}.conv5"]
            if j < 2:
                plan.conv(conv_args(cur.ptr(0), B, h, w, cur.stride, nf + 4 * g, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                    bias=bptr(c), s0=0.2, res1=t_cur.data_ptr(), res1_kind=L.SSR_F32, res1_stride=nf, s1=1.0,
                                    out=nxt.ptr(0), out_stride=nxt.stride,
                                    out32=t_nxt.data_ptr(), out32_mode=L.OUT32_NHWC, out32_stride=nf))
            else:
                t_blk = self.trunk_of(3 * blk)
                # (x5*0.2 + x_rdb3)*0.2 + x_rrdb
                plan.conv(conv_args(cur.ptr(0), B, h, w, cur.stride, nf + 4 * g, c.packed.data_ptr(), 3, c.cout, c.n_pad,
                                    bias=bptr(c), s0=0.04, res1=t_cur.data_ptr(), res1_kind=L.SSR_F32, res1_stride=nf, s1=0.2,
                                    res2=t_blk.data_ptr(), res2_kind=L.SSR_F32, res2_stride=nf, s2=1.0,
                                    out=nxt.ptr(0), out_stride=nxt.stride,
                                    out32=t_nxt.data_ptr(), out32_mode=L.OUT32_NHWC, out32_stride=nf))
        c = eng.cv["conv_body"]
        plan.conv(conv_args(self.body_out.ptr(), B, h, w, nf, nf, c.packed.data_ptr(), 3, c.cout, c.n_pad, bias=bptr(c),
                            res1=self.trunk_of(0).data_ptr(), res1_kind=L.SSR_F32, res1_stride=nf, s1=1.0,
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
