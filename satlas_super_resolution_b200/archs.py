"""Registry-facing network modules: SSR_RRDBNet and SSR_UNetDiscriminatorSN.

Same registry names, constructor signatures, forward contract (f32 NCHW in / out) and state_dict key schema as
/root/reference/ssr/archs/rrdbnet_arch.py:71-137 and discriminator_arch.py:11-71, so checkpoints and callers
(`build_network(opt['network_g'])`, ssr/utils/model_utils.py:17, ssr/infer.py:54) are interchangeable -- but `forward`
runs the tcgen05 engines of this package through the C ABI, not torch.nn.functional.  There is no CPU path:
calling forward on CPU tensors raises.
"""
import os
import weakref
from collections import OrderedDict

import torch
from torch import nn

from . import _lib as L
from . import weights
from .ops import FlatBuffer, cur_stream, lib
from .registry import ARCH_REGISTRY, _register


class _Token:
    """lifetime marker of one autograd context: a workspace is reusable once its token died or was consumed"""
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


class _FlatModule(nn.Module):
    """Keeps every learnable tensor of the module in ONE flat f32 buffer (+ one flat gradient buffer) once on the GPU."""

    _flat = None
    _flat_grad = None

    def _learnable_names(self):
        return [k for k, _ in self.named_parameters()]

    def _ensure_flat(self):
        params = OrderedDict(self.named_parameters())
        first = next(iter(params.values()))
        if not first.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: the B200 engine has no CPU path -- move the module to a CUDA device")
        if self._flat is not None and all(p.data_ptr() == self._flat.view(k).data_ptr() for k, p in params.items()):
            return False
        flat = FlatBuffer(OrderedDict((k, tuple(p.shape)) for k, p in params.items()), first.device)
        for k, p in params.items():
            flat.view(k).copy_(p.data)
            p.data = flat.view(k)
        self._flat = flat
        self._flat_grad = flat.like()
        self._engine = None
        return True

    def adopt(self, flat, flat_grad):
        """share the flat buffers of an ESRGANTrainer (models.SSRESRGANModel)"""
        for k, p in self.named_parameters():
            p.data = flat.view(k)
        self._flat, self._flat_grad = flat, flat_grad
        self._engine = None

    def _attach_grads(self):
        """make param.grad views of the flat gradient buffer (zeroing it when the optimizer dropped the grads)"""
        params = list(self.named_parameters())
        if all(p.grad is None for _, p in params):
            self._flat_grad.flat.zero_()
        for k, p in params:
            if p.grad is None or p.grad.data_ptr() != self._flat_grad.view(k).data_ptr():
                if p.grad is not None:
                    self._flat_grad.view(k).copy_(p.grad)
                p.grad = self._flat_grad.view(k)

    def grad_views(self):
        return {k: self._flat_grad.view(k) for k in self._flat.offsets}


def _conv(cout, cin, k=3, bias=True):
    m = nn.Module()
    m.weight = nn.Parameter(torch.empty(cout, cin, k, k))
    if bias:
        m.bias = nn.Parameter(torch.empty(cout))
    return m


# ======================================================================================= generator
class _RDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch):
        super().__init__()
        for k in range(1, 5):
            setattr(self, f"conv{k}", _conv(num_grow_ch, num_feat + (k - 1) * num_grow_ch))
        self.conv5 = _conv(num_feat, num_feat + 4 * num_grow_ch)


class _RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch):
        super().__init__()
        self.rdb1, self.rdb2, self.rdb3 = (_RDB(num_feat, num_grow_ch) for _ in range(3))


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, mod, grad_mode):
        eng = mod._get_engine()
        B, _, h, w = x.shape
        need_grad = grad_mode and anchor.requires_grad      # (grad mode is always off inside Function.forward)
        ctx.mod, ctx.shape, ctx.need_grad = mod, (B, h, w), need_grad
        if mod._weights_dirty():
            eng.repack()
        if need_grad:
            ws, token = mod._acquire(B, h // eng.unshuffle, w // eng.unshuffle)
            ctx.token, ctx.ws = token, ws
            out = eng.forward(x.contiguous(), train=True, ws=ws)
        else:
            out = eng.forward(x.contiguous(), train=False)
        return out.clone()

    @staticmethod
    def backward(ctx, d_out):
        mod = ctx.mod
        if ctx.need_grad:
            mod._attach_grads()
            B, h, w = ctx.shape
            mod._get_engine().backward(d_out.contiguous(), B, h, w, ws=ctx.ws)
            ctx.token.done = True
        return None, None, None, None


class SSR_RRDBNet(_FlatModule):
    """ESRGAN generator: conv_first -> num_block RRDBs -> conv_body (+skip) -> 2x nearest + conv (x log2(scale))
    -> conv_hr -> conv_last.  Registry name and signature of rrdbnet_arch.py:71-92."""

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale, self.num_in_ch, self.num_out_ch = scale, num_in_ch, num_out_ch
        self.num_feat, self.num_block, self.num_grow_ch = num_feat, num_block, num_grow_ch
        cin = num_in_ch * (4 if scale == 2 else 16 if scale == 1 else 1)
        self.conv_first = _conv(num_feat, cin)
        self.body = nn.Sequential(*[_RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = _conv(num_feat, num_feat)
        self.conv_up1 = _conv(num_feat, num_feat)
        self.conv_up2 = _conv(num_feat, num_feat)
        if scale in (8, 16):
            self.conv_up3 = _conv(num_feat, num_feat)
            if scale == 16:
                self.conv_up4 = _conv(num_feat, num_feat)
        self.conv_hr = _conv(num_feat, num_feat)
        self.conv_last = _conv(num_out_ch, num_feat)
        sd = weights.rrdbnet_state(num_in_ch, num_out_ch, scale, num_feat, num_block, num_grow_ch,
                                   seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        with torch.no_grad():
            for k, p in self.named_parameters():
                p.copy_(sd[k])
        self._engine = None
        self._anchor = torch.zeros(1, requires_grad=True)
        self._packed_version = None
        self._pool = {}

    # ---- engine plumbing
    def _get_engine(self):
        self._ensure_flat()
        if self._engine is None:
            from .generator import RRDBNetEngine
            self._engine = RRDBNetEngine(self._flat.views(), self.num_in_ch, self.num_out_ch, scale=self.scale, num_feat=self.num_feat,
                                         num_block=self.num_block, num_grow_ch=self.num_grow_ch, want_grad=True,
                                         grads=self._flat_grad.views())
            self._packed_version = None
            self._pool = {}
        return self._engine

    def _weights_dirty(self):
        ver = (sum(p._version for p in self.parameters()), getattr(self, "_external_version", 0))
        if ver != self._packed_version:
            self._packed_version = ver
            return True
        return False

    def mark_weights_changed(self):
        """call after modifying the flat parameter buffer outside torch (fused optimizer kernels)"""
        self._external_version = getattr(self, "_external_version", 0) + 1

    def _acquire(self, B, h, w):
        from .generator import _Workspace
        pool = self._pool.setdefault((B, h, w), [])
        for ws, ref in pool:
            tok = ref[0]() if ref[0] is not None else None
            if tok is None or tok.done:
                token = _Token()
                ref[0] = weakref.ref(token)
                return ws, token
        ws = _Workspace(self._engine, B, h, w, True)
        token = _Token()
        pool.append((ws, [weakref.ref(token)]))
        return ws, token

    def forward_split_bf16(self, x):
        """The tight-parity forward (tight.SplitBf16RRDBNet): operands as (hi, lo) bf16 pairs, ~2^-16 per layer -- a validation mode
        that agrees with the fp32 reference to ~1e-5; no autograd.  The production forward is `forward`."""
        if not x.is_cuda:
            raise RuntimeError("SSR_RRDBNet: the B200 engine has no CPU path (input must be a CUDA tensor)")
        from .tight import SplitBf16RRDBNet
        net = SplitBf16RRDBNet({k: v.detach() for k, v in self.state_dict().items()}, self.num_in_ch, self.num_out_ch, scale=self.scale,
                               num_feat=self.num_feat, num_block=self.num_block, num_grow_ch=self.num_grow_ch, device=x.device)
        return net.forward(x.float())

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("SSR_RRDBNet: the B200 engine has no CPU path (input must be a CUDA tensor)")
        if os.environ.get("SSR_PRECISION") == "split_bf16" and not torch.is_grad_enabled():
            return self.forward_split_bf16(x)      # ssr/infer.py / infer_grid.py under torch.no_grad(): tight-parity evaluation
        if x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("SSR_RRDBNet: the gradient w.r.t. the low-res input is never needed on the path and is not built")
        anchor = self._anchor if any(p.requires_grad for p in self.parameters()) else self._anchor.detach()
        return _GeneratorFn.apply(x.float(), anchor, self, torch.is_grad_enabled())


# ======================================================================================= discriminator
class _SNConv(nn.Module):
    """parameter / buffer names of torch.nn.utils.spectral_norm(nn.Conv2d(..., bias=False)) (legacy hook API)"""

    def __init__(self, cout, cin, k):
        super().__init__()
        self.weight_orig = nn.Parameter(torch.empty(cout, cin, k, k))
        self.register_buffer("weight_u", torch.empty(cout))
        self.register_buffer("weight_v", torch.empty(cin * k * k))


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, mod, grad_mode):
        eng = mod._get_engine()
        B, Cc, H, W = x.shape
        need = grad_mode and (anchor.requires_grad or x.requires_grad)
        ws, token = mod._acquire(B, H, W) if need else (eng.workspace(B, H, W), None)
        s = cur_stream()
        L.check(lib().ssr_ingest_nchw(x.contiguous().data_ptr(), L.SSR_F32, ws.x_in.ptr(), ws.x_in.stride, B, Cc, H, W,
                                      eng.cin_pad, 1.0, None, None, s))
        logits = eng.forward(ws, training=mod.training, stream=s)
        # sigma, u, v and the packed W / sigma operands are engine state that every forward overwrites (torch's spectral_norm
        # clones them per forward): remember which forward this graph belongs to
        mod._fwd_serial = getattr(mod, "_fwd_serial", 0) + 1
        ctx.serial = mod._fwd_serial
        ctx.mod, ctx.ws, ctx.token = mod, ws, token
        ctx.wgrad, ctx.dinput, ctx.x_shape = anchor.requires_grad, x.requires_grad, (B, Cc, H, W)
        return logits.clone()

    @staticmethod
    def backward(ctx, d_logits):
        mod, ws = ctx.mod, ctx.ws
        eng = mod._get_engine()
        if ctx.serial != mod._fwd_serial:
            raise RuntimeError("SSR_UNetDiscriminatorSN: backward of a forward that is no longer the latest one -- the spectral-norm "
                               "state (sigma, u, v, packed W / sigma) has been overwritten by a newer forward.  Call backward() after "
                               "each forward (as ssr_esrgan_model.py:215-227 does), not after several.")
        if ctx.wgrad:
            mod._attach_grads()
        eng.backward(ws, d_logits.contiguous(), need_wgrad=ctx.wgrad, need_dinput=ctx.dinput)
        dx = None
        if ctx.dinput:
            B, Cc, H, W = ctx.x_shape
            dx = torch.empty((B, Cc, H, W), dtype=torch.float32, device=d_logits.device)
            L.check(lib().ssr_egress_nchw(ws.d_in.ptr(), ws.d_in.stride, dx.data_ptr(), B, Cc, H, W, 1.0, 0, None, cur_stream()))
        if ctx.token is not None:
            ctx.token.done = True
        return dx, None, None, None


class SSR_UNetDiscriminatorSN(_FlatModule):
    """U-Net discriminator with spectral norm: registry name and signature of discriminator_arch.py:11-40."""

    def __init__(self, num_in_ch, num_feat=64, skip_connection=True):
        super().__init__()
        self.num_in_ch, self.num_feat, self.skip_connection = num_in_ch, num_feat, skip_connection
        nf = num_feat
        self.conv0 = _conv(nf, num_in_ch)
        self.conv1, self.conv2, self.conv3 = _SNConv(nf * 2, nf, 4), _SNConv(nf * 4, nf * 2, 4), _SNConv(nf * 8, nf * 4, 4)
        self.conv4, self.conv5, self.conv6 = _SNConv(nf * 4, nf * 8, 3), _SNConv(nf * 2, nf * 4, 3), _SNConv(nf, nf * 2, 3)
        self.conv7, self.conv8 = _SNConv(nf, nf, 3), _SNConv(nf, nf, 3)
        self.conv9 = _conv(1, nf)
        sd = weights.unet_disc_state(num_in_ch, num_feat, seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
        with torch.no_grad():
            own = self.state_dict()
            for k in own:
                own[k].copy_(sd[k])
        self._engine = None
        self._anchor = torch.zeros(1, requires_grad=True)
        self._pool = {}

    def _get_engine(self):
        changed = self._ensure_flat()
        bufs = dict(self.named_buffers())
        if self._engine is not None and any(self._engine.p[k].data_ptr() != b.data_ptr() for k, b in bufs.items()):
            changed = True
        if self._engine is None or changed:
            from .discriminator import UNetDiscEngine
            params = dict(self._flat.views())
            for k, b in bufs.items():
                if not b.is_contiguous():
                    b.data = b.data.contiguous()
                params[k] = b
            self._engine = UNetDiscEngine(params, self.num_in_ch, self.num_feat, self.skip_connection,
                                          grads=self._flat_grad.views())
            self._pool = {}
        return self._engine

    def _acquire(self, B, H, W):
        from .discriminator import _DWorkspace
        pool = self._pool.setdefault((B, H, W), [])
        for ws, ref in pool:
            tok = ref[0]() if ref[0] is not None else None
            if tok is None or tok.done:
                token = _Token()
                ref[0] = weakref.ref(token)
                return ws, token
        ws = _DWorkspace(self._engine, B, H, W)
        token = _Token()
        pool.append((ws, [weakref.ref(token)]))
        return ws, token

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("SSR_UNetDiscriminatorSN: the B200 engine has no CPU path (input must be a CUDA tensor)")
        anchor = self._anchor if any(p.requires_grad for p in self.parameters()) else self._anchor.detach()
        return _DiscFn.apply(x.float(), anchor, self, torch.is_grad_enabled())


_register(ARCH_REGISTRY, SSR_RRDBNet)
_register(ARCH_REGISTRY, SSR_UNetDiscriminatorSN)
