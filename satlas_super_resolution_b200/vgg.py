"""VGG19 feature extractor + perceptual (feature L1) loss engine -- basicsr PerceptualLoss as configured at
/root/reference/ssr/options/esrgan_s2naip_urban.yml:123-137 and called at ssr/models/ssr_esrgan_model.py:154.

The generated image and the ground truth go through the frozen network as ONE batch of 2B images (same weights, half
the launches); features are the PRE-ReLU conv outputs conv{b}_{last}; the backward pass runs only on the generated
half (B images): feature-L1 gradient -> conv^T chain with ReLU masks -> max-pool routing -> ... -> d(image)/std.
"""
import torch

from . import _lib as L
from .ops import Act, PackedConv, Packer, Plan, conv_args, cur_stream, lib

VGG19_CFG = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool1",
             ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool2",
             ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), "pool3",
             ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512), "pool4",
             ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512)]
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class PerceptualEngine:
    def __init__(self, params, layer_weights, perceptual_weight=1.0, use_input_norm=True, range_norm=False, split_upto=None):
        """params: 'conv1_1.weight' / '.bias' ... cuda f32 tensors (torchvision vgg19.features order).
        split_upto: with the side lane (forward_gt) the layers up to and including this pool run once per half of the 2B batch --
        the ground-truth half beside the generator forward -- and the deeper ones as ONE 2B batch after the join: at 16 x 16 and
        8 x 8 pixels a half batch fills 43 .. 86 % of the SMs.  'all' = every layer per half; default $SSR_VGG_SPLIT or pool3."""
        import os
        self.split_upto = split_upto or os.environ.get("SSR_VGG_SPLIT", "pool3")
        self.device = params["conv1_1.weight"].device
        dev = self.device
        self.layer_weights = dict(layer_weights)
        self.pw = perceptual_weight
        self.range_norm = range_norm
        names = [c[0] for c in VGG19_CFG if not isinstance(c, str)]
        last = max(names.index(k) for k in self.layer_weights)
        self.cfg = []
        for c in VGG19_CFG:
            self.cfg.append(c)
            if not isinstance(c, str) and names.index(c[0]) == last:
                break
        for k in self.layer_weights:
            idx = self.cfg.index(next(c for c in self.cfg if not isinstance(c, str) and c[0] == k))
            if idx + 1 < len(self.cfg) and not isinstance(self.cfg[idx + 1], str):
                raise NotImplementedError(f"perceptual layer {k}: only conv layers followed by a pool (or the last layer) "
                                          "are built as feature taps")
        for idx, c in enumerate(self.cfg):
            if isinstance(c, str) and self.cfg[idx - 1][0] not in self.layer_weights:
                raise NotImplementedError(f"{c}: a pool whose input conv is not a perceptual layer is not built "
                                          "(the shipped config taps conv1_2, 2_2, 3_4, 4_4, 5_4)")
        self.cv = {}
        for c in self.cfg:
            if isinstance(c, str):
                continue
            name, cin, cout = c
            self.cv[name] = PackedConv(params[f"{name}.weight"], params[f"{name}.bias"], 16 if cin == 3 else cin, True, dev)
        Packer(list(self.cv.values()), dev).run()
        if use_input_norm:
            self.mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32, device=dev)
            self.inv_std = 1.0 / torch.tensor(IMAGENET_STD, dtype=torch.float32, device=dev)
        else:
            self.mean = torch.zeros(3, dtype=torch.float32, device=dev)
            self.inv_std = torch.ones(3, dtype=torch.float32, device=dev)
        self._ws = {}

    def workspace(self, B, H, W):
        key = (B, H, W)
        if key not in self._ws:
            self._ws[key] = _PWorkspace(self, B, H, W)
        return self._ws[key]

    def _norm(self, ws):
        """(x + 1) / 2 of range_norm folded into the mean shift: v = x*0.5 + 0.5"""
        if not self.range_norm:
            return self.mean, self.inv_std
        ws.mean_t.copy_((self.mean - 0.5) / 0.5)
        ws.inv_std_t.copy_(self.inv_std * 0.5)
        return ws.mean_t, ws.inv_std_t

    def forward_gt(self, gt, lane):
        """The ground-truth half of the feature pass on its own (SSR_OVERLAP: issued on the side stream `lane` -- ops.SideLane,
        already forked -- while the generator's dense-block launches leave 20 SMs idle); loss_and_grad(..., gt_lane=lane) then runs
        the generated half only and joins the lane before the feature-L1 kernels."""
        B, _, H, W = gt.shape
        ws = self.workspace(B, H, W)
        assert not self.range_norm, "forward_gt: range_norm rewrites the normalisation constants on the main stream"
        mean, inv_std = self._norm(ws)
        L.check(lib().ssr_ingest_nchw(gt.data_ptr(), L.SSR_F32, ws.inp.ptr() + 2 * B * H * W * 16, 16, B, 3, H, W, 16, 1.0,
                                      mean.data_ptr(), inv_std.data_ptr(), lane.handle))
        ws.half_plans(self.split_upto)[1].run(lane.handle)

    def loss_and_grad(self, x, gt, loss_out, d_x, stream=None, gt_lane=None):
        """x, gt: f32 NCHW [B,3,H,W].  Adds the weighted perceptual loss to the device scalar `loss_out` and, when d_x is
        not None, ACCUMULATES d loss / d x into d_x (f32 NCHW).  gt_lane: forward_gt(gt, gt_lane) has been issued."""
        s = stream if stream is not None else cur_stream()
        B, _, H, W = x.shape
        ws = self.workspace(B, H, W)
        lb = lib()
        mean, inv_std = self._norm(ws)
        L.check(lb.ssr_ingest_nchw(x.data_ptr(), L.SSR_F32, ws.inp.ptr(), 16, B, 3, H, W, 16, 1.0, mean.data_ptr(),
                                   inv_std.data_ptr(), s))
        ws.loss_ptr[0] = loss_out.data_ptr()
        if gt_lane is not None:
            halves = ws.half_plans(self.split_upto)
            halves[0].run(s)
            gt_lane.join(s)
            halves[2].run(s)            # the layers below the split as one 2B batch (empty when every layer is split)
        else:
            L.check(lb.ssr_ingest_nchw(gt.data_ptr(), L.SSR_F32, ws.inp.ptr() + 2 * B * H * W * 16, 16, B, 3, H, W, 16, 1.0,
                                       mean.data_ptr(), inv_std.data_ptr(), s))
            ws.fwd.run(s)
        for fn, args in ws.loss_calls:
            L.check(fn(*args, loss_out.data_ptr(), s))
        if d_x is not None:
            ws.bwd.run(s)
            L.check(lb.ssr_egress_nchw(ws.d_inp.ptr(), 16, d_x.data_ptr(), B, 3, H, W, 1.0, 1, inv_std.data_ptr(), s))


class _PWorkspace:
    def __init__(self, eng, B, H, W):
        dev = eng.device
        self.B, self.H, self.W = B, H, W
        B2 = 2 * B
        self.mean_t = torch.zeros(3, dtype=torch.float32, device=dev)
        self.inv_std_t = torch.ones(3, dtype=torch.float32, device=dev)
        self.loss_ptr = [0]
        self.inp = Act(B2, H, W, 16, dev)
        fwd, bwd = Plan(), Plan()
        self.loss_calls = []
        lb = lib()
        cur, hh, ww = self.inp, H, W
        cur_c = 16
        acts = {}        # conv name -> (output Act, hh, ww, is_feature)
        order = []
        for idx, c in enumerate(eng.cfg):
            if isinstance(c, str):
                prev_name = order[-1][1]
                src = acts[prev_name][0]
                dst = Act(B2, hh // 2, ww // 2, src.C, dev)
                fwd.add(lb.ssr_maxpool_relu, src.ptr(), dst.ptr(), B2, hh, ww, src.C)
                order.append(("pool", c, src, dst, hh, ww))
                hh, ww = hh // 2, ww // 2
                cur, cur_c = dst, dst.C
                continue
            name, cin, cout = c
            pc = eng.cv[name]
            is_feat = name in eng.layer_weights
            out = Act(B2, hh, ww, cout, dev)
            fwd.conv(conv_args(cur.ptr(), B2, hh, ww, cur.stride, cur_c, pc.packed.data_ptr(), 3, cout, pc.n_pad,
                               bias=pc.bias.data_ptr(), act=0 if is_feat else 2, out=out.ptr(), out_stride=cout))
            acts[name] = (out, hh, ww, is_feat)
            order.append(("conv", name, cur, out, hh, ww, cur_c))
            if is_feat:
                n_half = B * hh * ww * cout
                self.loss_calls.append((lb.ssr_feat_l1, (out.ptr(), n_half, eng.layer_weights[name] * eng.pw / n_half)))
            cur, cur_c = out, cout
        self.fwd = fwd
        self.order = order
        self._eng = eng
        self._halves = None
        # ---------------- backward over the generated half (first B images of every buffer)
        self.d_inp = Act(B, H, W, 16, dev)
        self._keep = []
        g = None          # gradient flowing into the current position (w.r.t. the OUTPUT of order[i])
        for i in range(len(order) - 1, -1, -1):
            item = order[i]
            if item[0] == "pool":
                # g is d(pool out); routed to the feature layer below by feat_grad (handled at that conv)
                continue
            _, name, src, out, hh, ww, cin_c = item
            pc = eng.cv[name]
            cout = out.C
            is_feat = acts[name][3]
            if is_feat:
                dF = Act(B, hh, ww, cout, dev)
                self._keep.append(dF)
                n_half = B * hh * ww * cout
                bwd.add(lb.ssr_feat_grad, out.ptr(), g.ptr() if g is not None else None, dF.ptr(), B, hh, ww, cout,
                        eng.layer_weights[name] * eng.pw / n_half)
                dy = dF
            else:
                dy = g   # already masked by this layer's ReLU in the producer's epilogue
            # input gradient of this conv
            if i == 0:
                bwd.conv(conv_args(dy.ptr(), B, hh, ww, cout, cout, pc.packed_dg.data_ptr(), 3, 3, pc.n_pad_dg,
                                   out=self.d_inp.ptr(), out_stride=16))
                break
            below = order[i - 1]
            dx = Act(B, hh, ww, cin_c, dev)
            self._keep.append(dx)
            if below[0] == "conv" and not acts[below[1]][3]:
                # the input is relu(conv_below): mask with its stored (post-ReLU) output
                bwd.conv(conv_args(dy.ptr(), B, hh, ww, cout, cout, pc.packed_dg.data_ptr(), 3, cin_c, pc.n_pad_dg,
                                   mask=below[3].ptr(), mask_stride=cin_c, mask_lo=0, mask_relu=1,
                                   out=dx.ptr(), out_stride=cin_c))
            else:
                # the input is a pool output (routing + ReLU handled by feat_grad of the feature layer below the pool)
                bwd.conv(conv_args(dy.ptr(), B, hh, ww, cout, cout, pc.packed_dg.data_ptr(), 3, cin_c, pc.n_pad_dg,
                                   out=dx.ptr(), out_stride=cin_c))
            g = dx
        self.bwd = bwd

    def half_plans(self, split_upto="all"):
        """[plan over the generated half, plan over the ground-truth half, plan over the whole 2B batch]: the layers up to and
        including pool `split_upto` once per half, the deeper ones as one batch -- same buffers, same launches as self.fwd otherwise"""
        if self._halves is None:
            lb = lib()
            B = self.B
            names = [item[1] for item in self.order]
            n_split = names.index(split_upto) + 1 if split_upto in names else len(self.order)
            off = lambda a, half: a.ptr() + half * B * a.H * a.W * a.C * 2

            def emit(plan, item, half, nb):
                if item[0] == "pool":
                    _, _, src, dst, hh, ww = item
                    plan.add(lb.ssr_maxpool_relu, off(src, half), off(dst, half), nb, hh, ww, src.C)
                    return
                _, name, src, out, hh, ww, cin_c = item
                pc = self._eng.cv[name]
                plan.conv(conv_args(off(src, half), nb, hh, ww, src.stride, cin_c, pc.packed.data_ptr(), 3, out.C, pc.n_pad,
                                    bias=pc.bias.data_ptr(), act=0 if name in self._eng.layer_weights else 2,
                                    out=off(out, half), out_stride=out.C))

            plans = [Plan(), Plan(), Plan()]
            for half in (0, 1):
                for item in self.order[:n_split]:
                    emit(plans[half], item, half, B)
            for item in self.order[n_split:]:
                emit(plans[2], item, 0, 2 * B)
            self._halves = plans
        return self._halves
