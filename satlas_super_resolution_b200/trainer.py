"""The ESRGAN training step (feed_data + optimize_parameters) driven entirely through libssr_b200.

Statement order, loss weights, detach points and the three discriminator forwards follow
/root/reference/ssr/models/ssr_esrgan_model.py:104-233; what differs is HOW each statement executes:
every tensor op is a kernel of the C ABI, parameters / gradients / Adam moments / EMA live in flat f32 buffers
(one fused Adam(+EMA) launch and one all-reduce per network), and the whole step is replayable as one CUDA graph.
"""
import math
from collections import OrderedDict

import torch

from . import _lib as L
from .discriminator import UNetDiscEngine
from .generator import RRDBNetEngine
from .ops import FlatBuffer, SideLane, allreduce_sum_, cur_stream, lib, overlap_enabled
from .vgg import PerceptualEngine

LOSS_KEYS = ["l_g_pix", "l_g_percep", "l_g_gan", "l_d_real", "out_d_real", "l_d_fake", "out_d_fake", "l_g_ssim"]   # slots of loss_dev
LOG_ORDER = ["l_g_pix", "l_g_percep", "l_g_ssim", "l_g_gan", "l_d_real", "out_d_real", "l_d_fake", "out_d_fake"]     # ssr_esrgan_model.py:149-226


def gaussian_taps(ksize=51, sigma=0.0):
    """cv2.getGaussianKernel(ksize, sigma) (basicsr USMSharp): sigma <= 0 -> 0.3*((ksize-1)*0.5 - 1) + 0.8."""
    if sigma <= 0:
        sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    c = (ksize - 1) / 2
    k = [math.exp(-((i - c) ** 2) / (2 * sigma * sigma)) for i in range(ksize)]
    s = sum(k)
    return [v / s for v in k]


class FusedAdamState:
    """Adam moments (and the EMA copy) for one flat parameter buffer; one ssr_adam_ema launch per step."""

    def __init__(self, params: FlatBuffer, grads: FlatBuffer, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0,
                 ema: FlatBuffer = None, ema_decay=0.0):
        self.p, self.g = params, grads
        self.m, self.v = params.like(), params.like()
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.ema, self.ema_decay = ema, ema_decay
        self.step_count = 0
        # device-resident [lr, 1-b1^t, sqrt(1-b2^t), t, b1, b2] for CUDA-graph replays (advanced by ssr_adam_tick in-graph)
        self.hyper_dev = torch.zeros(6, dtype=torch.float32, device=params.flat.device)
        self._dev_lr = None

    def sync_device_hyper(self):
        """make the device block agree with the host state (called before capturing / when lr or the step changed eagerly)"""
        self.hyper_dev.copy_(torch.tensor([self.lr, 0.0, 0.0, float(self.step_count), self.betas[0], self.betas[1]]))
        self._dev_lr = self.lr

    def before_replay(self):
        self.step_count += 1
        if self._dev_lr != self.lr:
            self.hyper_dev[0:1].fill_(self.lr)
            self._dev_lr = self.lr

    def step(self, grad_scale=1.0, stream=None, from_device=False):
        s_ = stream if stream is not None else cur_stream()
        if from_device:
            L.check(lib().ssr_adam_tick(self.hyper_dev.data_ptr(), s_))
        else:
            self.step_count += 1
        L.check(lib().ssr_adam_ema(self.p.flat.data_ptr(), self.g.flat.data_ptr(), self.m.flat.data_ptr(),
                                   self.v.flat.data_ptr(), self.ema.flat.data_ptr() if self.ema is not None else None,
                                   self.p.numel, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                   self.step_count, self.ema_decay, grad_scale,
                                   self.hyper_dev.data_ptr() if from_device else None,
                                   stream if stream is not None else cur_stream()))


class ESRGANTrainer:
    D_BUFFERS = ("weight_u", "weight_v")

    def __init__(self, g_state, d_state, vgg_state, cfg=None, device="cuda", process_group=None):
        cfg = dict(cfg or {})
        self.cfg = cfg
        self.device = torch.device(device)
        dev = self.device
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        net_g = cfg.get("network_g", {})
        self.num_in_ch = net_g.get("num_in_ch", g_state["conv_first.weight"].shape[1])
        self.num_block = net_g.get("num_block", 23)
        self.scale = net_g.get("scale", 4)
        # ---- generator: flat params / grads / EMA
        self.gbuf = FlatBuffer(OrderedDict((k, tuple(v.shape)) for k, v in g_state.items()), dev)
        for k, v in g_state.items():
            self.gbuf.view(k).copy_(v)
        self.ggrad = self.gbuf.like()
        self.ema_decay = cfg.get("ema_decay", 0.999)
        self.gema = self.gbuf.like() if self.ema_decay > 0 else None
        if self.gema is not None:
            self.gema.flat.copy_(self.gbuf.flat)          # model_ema(0): ssr_esrgan_model.py:49
        self.G = RRDBNetEngine(self.gbuf.views(), self.num_in_ch, g_state["conv_last.weight"].shape[0], scale=self.scale,
                               num_feat=g_state["conv_first.weight"].shape[0], num_block=self.num_block,
                               num_grow_ch=g_state["body.0.rdb1.conv1.weight"].shape[0], want_grad=True,
                               grads=self.ggrad.views(), overlap=cfg.get("overlap"))
        self.G_ema = None
        # ---- discriminator
        learn = OrderedDict((k, tuple(v.shape)) for k, v in d_state.items() if not k.endswith(self.D_BUFFERS))
        self.dbuf = FlatBuffer(learn, dev)
        for k in learn:
            self.dbuf.view(k).copy_(d_state[k])
        self.dgrad = self.dbuf.like()
        self.d_uv = {k: v.detach().clone().to(dev, torch.float32) for k, v in d_state.items() if k.endswith(self.D_BUFFERS)}
        self.d_in_ch = d_state["conv0.weight"].shape[1]
        d_params = dict(self.dbuf.views())
        d_params.update(self.d_uv)
        self.D = UNetDiscEngine(d_params, self.d_in_ch, num_feat=d_state["conv0.weight"].shape[0], grads=self.dgrad.views())
        # ---- losses
        self.pixel_weight = cfg.get("pixel_weight", 1.0)
        self.gan_weight = cfg.get("gan_weight", 0.1)
        self.ssim_weight = cfg.get("ssim_weight", 0.0)   # train.ssim_opt (ssr_esrgan_model.py:87-90, 163-166); 0 = no SSIM term
        self.P = None
        if vgg_state is not None and cfg.get("perceptual", True):
            vp = {k: v.to(dev, torch.float32).contiguous() for k, v in vgg_state.items()}
            self.P = PerceptualEngine(vp, cfg.get("layer_weights", {"conv1_2": 0.1, "conv2_2": 0.1, "conv3_4": 1.0,
                                                                     "conv4_4": 1.0, "conv5_4": 1.0}),
                                      cfg.get("perceptual_weight", 1.0), cfg.get("use_input_norm", True),
                                      cfg.get("range_norm", False), split_upto=cfg.get("vgg_split"))
        self.feed_disc_lr = cfg.get("feed_disc_lr", True)
        self.old_hr_ch = 0   # set by feed_data when the batch carries an `old_hr` image (ssr_esrgan_model.py:112-114)
        self.l1_gt_usm = cfg.get("l1_gt_usm", True)
        self.percep_gt_usm = cfg.get("percep_gt_usm", True)
        self.gan_gt_usm = cfg.get("gan_gt_usm", False)
        self.net_d_iters = cfg.get("net_d_iters", 1)
        self.net_d_init_iters = cfg.get("net_d_init_iters", 0)
        lr = cfg.get("lr", 1e-4)
        betas = tuple(cfg.get("betas", (0.9, 0.99)))
        # train.optim_d has its own lr / betas / weight_decay (esrgan_s2naip_urban.yml:103-107); default = optim_g's
        self.opt_g = FusedAdamState(self.gbuf, self.ggrad, lr, betas, weight_decay=cfg.get("weight_decay", 0.0),
                                    ema=self.gema, ema_decay=self.ema_decay)
        self.opt_d = FusedAdamState(self.dbuf, self.dgrad, cfg.get("lr_d", lr), tuple(cfg.get("betas_d", betas)),
                                    weight_decay=cfg.get("weight_decay_d", cfg.get("weight_decay", 0.0)))
        if self.pg is not None:
            self.sync_replicas()
        self.loss_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        self.usm_taps = (C_float_array(gaussian_taps(51, 0)), 51)
        self._io = {}
        self._graphs = {}
        self._warm = set()
        self._last_mode = "eager"
        self.use_graph = bool(cfg.get("cuda_graph", False))
        # the side lane (ops.overlap_enabled): the ground-truth half of the VGG pass and the discriminator's weight preparation
        # run on a side stream beside the generator's dense-block launches (128 of the 148 SMs), like the dense blocks'
        # weight gradients beside the next input-gradient launch (generator._build_backward)
        self.overlap = overlap_enabled("fwd", cfg.get("overlap"))
        # one CUDA graph per phase also on one GPU (what world > 1 replays around its all-reduces): a test hook for that capture path
        self.phase_graphs = bool(cfg.get("phase_graphs", False))
        self._cap_stream = None
        self.log_dict = OrderedDict()

    def replicated_tensors(self):
        """everything DistributedDataParallel broadcasts from rank 0 at construction (parameters and buffers) plus the EMA copy"""
        out = [self.gbuf.flat, self.dbuf.flat] + [self.d_uv[k] for k in sorted(self.d_uv)]
        if self.gema is not None:
            out.append(self.gema.flat)
        return out

    def sync_replicas(self):
        """The reference seeds every rank with manual_seed + rank (ssr/utils/options.py:81) and relies on DDP broadcasting rank 0's
        parameters and buffers when the wrapper is built (basicsr model_to_device, ssr_esrgan_model.py:54): do the same, so
        replicas that were initialised differently start -- and, applying the same averaged gradients, stay -- identical."""
        from .ops import broadcast_from_rank0_
        broadcast_from_rank0_(self.replicated_tensors(), self.pg)

    # ------------------------------------------------------------------ data
    def _io_for(self, B, C_lr, h, w, H, W):
        key = (B, C_lr, h, w, H, W)
        io = self._io.get(key)
        if io is None:
            dev = self.device
            io = dict(lr_u8=torch.empty((B, C_lr, h, w), dtype=torch.uint8, device=dev),
                      hr_u8=torch.empty((B, 3, H, W), dtype=torch.uint8, device=dev),
                      lr=torch.empty((B, C_lr, h, w), dtype=torch.float32, device=dev),
                      gt=torch.empty((B, 3, H, W), dtype=torch.float32, device=dev),
                      gt_usm=torch.empty((B, 3, H, W), dtype=torch.float32, device=dev),
                      old_u8=torch.empty((B, 3, H, W), dtype=torch.uint8, device=dev),
                      old_hr=torch.empty((B, 3, H, W), dtype=torch.float32, device=dev),
                      usm_scratch=torch.empty((3, B, 3, H, W), dtype=torch.float32, device=dev),
                      d_out=torch.empty((B, 3, H, W), dtype=torch.float32, device=dev),
                      ssim_scratch=(torch.empty((3, B, 3, H, W), dtype=torch.float32, device=dev) if self.ssim_weight else None),
                      d_logits=torch.empty((B, 1, H, W), dtype=torch.float32, device=dev))
            self._io[key] = io
        return io

    @torch.no_grad()
    def feed_data(self, lr_u8, hr_u8, old_hr_u8=None, stream=None):
        """ssr_esrgan_model.py:104-117: uint8 -> float / 255 (+ USM-sharpened ground truth, + the optional `old_hr` image the
        discriminator is conditioned on).  The uint8 tensors may be host (pinned) or device tensors; the H2D copy is part of
        this call."""
        s = stream if stream is not None else cur_stream()
        B, C_lr, h, w = lr_u8.shape
        H, W = hr_u8.shape[-2:]
        io = self._io_for(B, C_lr, h, w, H, W)
        io["lr_u8"].copy_(lr_u8, non_blocking=True)
        io["hr_u8"].copy_(hr_u8, non_blocking=True)
        old_ch = 0
        if old_hr_u8 is not None:
            io["old_u8"].copy_(old_hr_u8, non_blocking=True)
            old_ch = 3
        if old_ch != self.old_hr_ch:
            self.old_hr_ch = old_ch      # the recorded step changes shape: re-warm / re-capture
            self._warm.clear()
            self._graphs.clear()
        want = 3 + (C_lr if self.feed_disc_lr else 0) + old_ch
        if want != self.d_in_ch:
            raise ValueError(f"network_d.num_in_ch is {self.d_in_ch} but the discriminator input has {want} channels "
                             f"(3 image + {C_lr if self.feed_disc_lr else 0} low-res + {old_ch} old_hr)")
        self._feed_kernels(io, s)
        self.io = io
        self.lr, self.gt, self.gt_usm = io["lr"], io["gt"], io["gt_usm"]

    def _feed_kernels(self, io, s):
        lb = lib()
        L.check(lb.ssr_u8_to_f32(io["lr_u8"].data_ptr(), io["lr"].data_ptr(), io["lr"].numel(), 1.0 / 255.0, s))
        L.check(lb.ssr_u8_to_f32(io["hr_u8"].data_ptr(), io["gt"].data_ptr(), io["gt"].numel(), 1.0 / 255.0, s))
        if self.old_hr_ch:
            L.check(lb.ssr_u8_to_f32(io["old_u8"].data_ptr(), io["old_hr"].data_ptr(), io["old_hr"].numel(), 1.0 / 255.0, s))
        B, _, H, W = io["gt"].shape
        taps, n = self.usm_taps
        L.check(lb.ssr_usm_sharp(io["gt"].data_ptr(), io["gt_usm"].data_ptr(), io["usm_scratch"].data_ptr(), B * 3, H, W, taps, n,
                                 0.5, 10.0, s))

    # ------------------------------------------------------------------ the step
    def _step_phase(self, phase, io, do_g, s, graph_mode=False):
        """The step in four collective-free phases (each can be a CUDA graph even when NCCL runs between them):
        1 = G forward, generator losses, frozen-D pass, G backward          (ssr_esrgan_model.py:136-192)
        2 = D real pass, D fake pass                                         (:196-227)
        3 = Adam(G) + EMA                                                    (:193, :230-231)
        4 = Adam(D)                                                          (:228)
        The reference runs optimizer_g.step() (:193) before the discriminator passes; they read neither the generator's parameters
        nor its gradients (only self.output), so 3 commutes with 2 -- which lets the all-reduce of the G gradients run UNDER the
        discriminator passes and the one of the D gradients under Adam(G)."""
        lb = lib()
        lr, gt, gt_usm = io["lr"], io["gt"], io["gt_usm"]
        B, C_lr, h, w = lr.shape
        H, W = gt.shape[-2:]
        n_img = gt.numel()
        l1_gt = gt_usm if self.l1_gt_usm else gt
        percep_gt = gt_usm if self.percep_gt_usm else gt
        gan_gt = gt_usm if self.gan_gt_usm else gt
        loss = self.loss_dev
        lp = lambda i: loss.data_ptr() + 4 * i
        d_out, d_logits = io["d_out"], io["d_logits"]
        gws = self.G.workspace(B, h // self.G.unshuffle, w // self.G.unshuffle, True)
        dws = self.D.workspace(B, H, W)
        n_logit = B * H * W
        cl = C_lr if self.feed_disc_lr else 0
        f = H // h
        ce = self.old_hr_ch

        def disc_in(img):
            # torch.cat((img, lr_resized, old_hr), 1) -- ssr_esrgan_model.py:171-178, 202-213 -- written straight as NHWC bf16
            L.check(lb.ssr_disc_input_ex(img.data_ptr(), 3, gws.in0.ptr() if cl else None, gws.in0.stride, cl, f,
                                         io["old_hr"].data_ptr() if ce else None, ce, dws.x_in.ptr(), dws.x_in.stride, B, H, W, s))

        if phase == 1:
            loss.zero_()
            self.G.repack(s)
            lane = None
            if self.overlap and do_g and not (self.P is not None and self.P.range_norm):
                # nothing below depends on the generator: issue it beside the dense-block launches
                lane = SideLane.get(self.device)
                lane.fork(s)
                self.D.prepare_weights(True, lane.handle)
                if self.P is not None:
                    self.P.forward_gt(percep_gt, lane)
            out = self.G.forward(lr, train=True, stream=s)
            self.output = out
            if do_g:
                self.ggrad.flat.zero_()
                L.check(lb.ssr_l1_loss(out.data_ptr(), l1_gt.data_ptr(), n_img, self.pixel_weight, lp(0), d_out.data_ptr(), 0, s))
                if self.P is not None:
                    self.P.loss_and_grad(out, percep_gt, loss[1:2], d_out, s, gt_lane=lane)
                if self.ssim_weight:
                    # l_g_ssim = ssim_loss(self.output, percep_gt) -- ssr_esrgan_model.py:163-166
                    L.check(lb.ssr_ssim_loss(out.data_ptr(), percep_gt.data_ptr(), B * 3, H, W, self.ssim_weight, lp(7),
                                             d_out.data_ptr(), 1, io["ssim_scratch"].data_ptr(), s))
                disc_in(out)
                if lane is not None:
                    lane.join(s)
                logits = self.D.forward(dws, training=True, stream=s, prepared=lane is not None)
                L.check(lb.ssr_bce_logits(logits.data_ptr(), n_logit, 1.0, self.gan_weight, lp(2), None, d_logits.data_ptr(), s))
                self.D.backward(dws, d_logits, need_wgrad=False, need_dinput=True, stream=s)
                L.check(lb.ssr_egress_nchw(dws.d_in.ptr(), dws.d_in.stride, d_out.data_ptr(), B, 3, H, W, 1.0, 1, None, s))
                self.G.backward(d_out, B, h, w, s)
        elif phase == 2:
            out = self.output
            self.dgrad.flat.zero_()
            disc_in(gan_gt)
            logits = self.D.forward(dws, training=True, stream=s)
            L.check(lb.ssr_bce_logits(logits.data_ptr(), n_logit, 1.0, 1.0, lp(3), lp(4), d_logits.data_ptr(), s))
            self.D.backward(dws, d_logits, need_wgrad=True, stream=s)
            disc_in(out)
            logits = self.D.forward(dws, training=True, stream=s)
            L.check(lb.ssr_bce_logits(logits.data_ptr(), n_logit, 0.0, 1.0, lp(5), lp(6), d_logits.data_ptr(), s))
            self.D.backward(dws, d_logits, need_wgrad=True, stream=s)
        elif phase == 3:
            if do_g:
                self.opt_g.step(1.0 / self.world, s, from_device=graph_mode)
            elif self.gema is not None:
                # model_ema runs every iteration in the reference (:230-231), also when the generator step is skipped
                L.check(lb.ssr_ema_update(self.gema.flat.data_ptr(), self.gbuf.flat.data_ptr(), self.gbuf.numel, self.ema_decay, s))
        else:
            self.opt_d.step(1.0 / self.world, s, from_device=graph_mode)

    def _exchange_async(self, flat):
        """sum all-reduce of one flat gradient buffer on NCCL's own stream (ordered behind everything issued so far); the returned
        handle's wait() orders the current stream behind it -- the host never blocks"""
        return torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=self.pg, async_op=True)

    def _run_step(self, run_phase, do_g):
        """phases with the gradient exchanges hidden: G's all-reduce under the discriminator passes, D's under Adam(G)"""
        run_phase(1)
        w_g = self._exchange_async(self.ggrad.flat) if (self.world > 1 and do_g) else None
        run_phase(2)
        w_d = self._exchange_async(self.dgrad.flat) if self.world > 1 else None
        if w_g is not None:
            w_g.wait()
        run_phase(3)
        if w_d is not None:
            w_d.wait()
        run_phase(4)

    def _step_kernels(self, io, do_g, s, graph_mode=False):
        self._run_step(lambda ph: self._step_phase(ph, io, do_g, s, graph_mode), do_g)

    def optimize_parameters(self, current_iter=1):
        do_g = (current_iter % self.net_d_iters == 0) and (current_iter > self.net_d_init_iters)
        self._last_do_g = do_g
        key = (id(self.io), do_g)
        if not self.use_graph or key not in self._warm:
            # eager: also the mandatory first pass per shape (allocates workspaces, sets kernel attributes)
            self._step_kernels(self.io, do_g, cur_stream())
            self._warm.add(key)
            self._last_mode = "eager"
            return
        graphs = self._graphs.get(key)
        if graphs is None:
            # capture once: the whole step as ONE graph on a single GPU, one graph per phase around the NCCL calls otherwise
            torch.cuda.synchronize()
            graphs = []
            # with a side lane the capture stream gets a higher priority than the lane's: where both have thread blocks pending,
            # the main lane's (the dense-block clusters) are placed first and the side lane's fill what is left
            cap = {}
            if self.overlap or self.G.overlap:
                if self._cap_stream is None:
                    self._cap_stream = torch.cuda.Stream(device=self.device, priority=-1)
                cap = dict(stream=self._cap_stream)
            if self.world == 1 and not self.phase_graphs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, **cap):
                    self._step_kernels(self.io, do_g, cur_stream(), graph_mode=True)
                graphs.append(g)
            else:
                pool = None
                for phase in (1, 2, 3, 4):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool, **cap):
                        self._step_phase(phase, self.io, do_g, cur_stream(), graph_mode=True)
                    pool = g.pool()
                    graphs.append(g)
            self._graphs[key] = graphs
            self.opt_g.sync_device_hyper()
            self.opt_d.sync_device_hyper()
        elif self._last_mode != "graph":
            self.opt_g.sync_device_hyper()
            self.opt_d.sync_device_hyper()
        self._last_mode = "graph"
        if do_g:
            self.opt_g.before_replay()
        self.opt_d.before_replay()
        if len(graphs) == 1:
            graphs[0].replay()
        else:
            self._run_step(lambda ph: graphs[ph - 1].replay(), do_g)

    def get_current_log(self):
        """loss scalars (one D2H read, only when asked -- the reference syncs every iteration at :233)"""
        vals = self.loss_dev.tolist()
        log = OrderedDict()
        for k in LOG_ORDER:
            i = LOSS_KEYS.index(k)
            if (k == "l_g_percep" and self.P is None) or (k == "l_g_ssim" and not self.ssim_weight):
                continue
            if k.startswith("l_g_") and not getattr(self, "_last_do_g", True):
                continue      # no generator step this iteration (net_d_iters / net_d_init_iters): the reference logs no l_g_*
            log[k] = vals[i]
        if self.world > 1:
            t = torch.tensor(list(log.values()), device=self.device)
            torch.distributed.reduce(t, dst=0, group=self.pg)
            t = t / self.world
            log = OrderedDict(zip(log.keys(), t.tolist()))
        self.log_dict = log
        return log

    # ------------------------------------------------------------------ state access (reference key schema)
    def g_state_dict(self, ema=False):
        buf = self.gema if ema else self.gbuf
        return OrderedDict((k, buf.view(k).detach().clone()) for k in buf.offsets)

    def d_state_dict(self):
        sd = OrderedDict((k, self.dbuf.view(k).detach().clone()) for k in self.dbuf.offsets)
        sd.update({k: v.detach().clone() for k, v in self.d_uv.items()})
        return sd

    def g_grads(self):
        return OrderedDict((k, self.ggrad.view(k)) for k in self.ggrad.offsets)

    def d_grads(self):
        return OrderedDict((k, self.dgrad.view(k)) for k in self.dgrad.offsets)

    @torch.no_grad()
    def test(self, lr=None):
        """ssr_esrgan_model.py:235-244: EMA generator in eval mode."""
        if self.G_ema is None:
            src = self.gema if self.gema is not None else self.gbuf
            self.G_ema = RRDBNetEngine(src.views(), self.num_in_ch, self.G.cout, scale=self.scale, num_feat=self.G.nf,
                                       num_block=self.num_block, num_grow_ch=self.G.g, want_grad=False)
        self.G_ema.repack()
        x = lr if lr is not None else self.lr
        self.output = self.G_ema.forward(x.contiguous(), train=False).clone()
        return self.output


def C_float_array(vals):
    import ctypes
    return (ctypes.c_float * len(vals))(*vals)
