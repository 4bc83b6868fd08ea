"""Host-side helpers over the C ABI: device buffers, pre-built argument structs and launch plans.

torch is used for device memory and streams only; every arithmetic op below is a kernel of
libssr_b200.so called through ctypes (include/ssr_b200.h).
"""
import ctypes as C

import torch

from . import _lib as L
from ._protos import PackDesc

BF16 = torch.bfloat16


def lib():
    return L.load()


def cur_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(v, m):
    return (v + m - 1) // m * m


class Act:
    """A bf16 NHWC activation buffer [B, H, W, stride]; `ch(lo)` gives the device pointer of a channel slice."""

    def __init__(self, B, H, W, C_, device, zero=False):
        self.B, self.H, self.W, self.C = B, H, W, C_
        self.t = (torch.zeros if zero else torch.empty)((B, H, W, C_), dtype=BF16, device=device)

    @property
    def stride(self):
        return self.C

    def ptr(self, ch=0):
        return self.t.data_ptr() + 2 * ch


_FORK, _JOIN, _SIDE = "fork", "join", "side"     # lane markers inside Plan.calls (never equal to a ctypes function)


def overlap_enabled(part=None, setting=None):
    """The side lane: independent launches run on a second stream beside the resident dense-block launches,
    which occupy 128 of the 148 SMs (one 4-CTA cluster per image at B = 32).  Parts: 'bwd' = a group's dense-block weight
    gradients beside the next input-gradient launch; 'fwd' = the ground-truth half of the VGG pass and the discriminator's weight
    preparation beside the generator forward.  `setting` (an option value) or $SSR_OVERLAP: 1 / True = both parts (the default),
    0 / False = one stream, 'fwd' / 'bwd' = that part only.  Read when a plan / trainer is BUILT."""
    import os
    v = os.environ.get("SSR_OVERLAP", "1") if setting is None else setting
    if isinstance(v, bool):
        return v
    v = str(v).lower()
    if v in ("1", "true", "on", "all"):
        return True
    parts = {p.strip() for p in v.split(",")} & {"fwd", "bwd"}
    return bool(parts) if part is None else part in parts


class SideLane:
    """A second stream + fork / join events (capturable: inside a CUDA-graph capture the events become graph edges; every
    fork must be joined before the capture ends)."""

    _per_device = {}

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.handle = C.c_void_p(self.stream.cuda_stream)
        self.dirty = False

    @classmethod
    def get(cls, device=None):
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
        lane = cls._per_device.get(idx)
        if lane is None:
            lane = cls._per_device[idx] = SideLane(torch.device("cuda", idx))
        return lane

    @staticmethod
    def _main(s):
        cur = torch.cuda.current_stream()
        if s is None or (s.value or 0) == cur.cuda_stream:
            return cur
        return torch.cuda.ExternalStream(s.value)

    def fork(self, s=None):
        """the side stream waits for everything issued on the main stream `s` so far"""
        ev = torch.cuda.Event()
        ev.record(self._main(s))
        self.stream.wait_event(ev)
        self.dirty = True

    def join(self, s=None):
        """the main stream waits for everything issued on the side stream so far"""
        if not self.dirty:
            return
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self._main(s).wait_event(ev)


class Plan:
    """A recorded, allocation-free sequence of C-ABI calls (replayed every step, CUDA-graph capturable).

    Calls recorded inside `with plan.side():` go to the side stream (ops.SideLane); `fork()` makes the side stream wait for the
    main-lane calls recorded so far, `join()` the main lane for the side calls recorded so far.  `run` joins at its end."""

    def __init__(self):
        self.calls = []
        self.keep = []
        self._lane = 0
        self.has_side = False

    def _emit(self, fn, args):
        if self._lane:
            self.calls.append((_SIDE, fn, args))
            self.has_side = True
        else:
            self.calls.append((fn, args))

    def add(self, fn, *args):
        self._emit(fn, args)

    def conv(self, args):
        self.keep.append(args)
        self._emit(lib().ssr_conv_tc, (C.byref(args),))

    def chain(self, args_list):
        """consecutive convs over one image geometry, each reading what the previous ones wrote: ONE launch (ssr_conv_tc_chain)"""
        arr = (L.ConvTcArgs * len(args_list))(*args_list)
        self.keep.append(arr)
        self._emit(lib().ssr_conv_tc_chain, (arr, len(args_list)))

    def chain_acc(self, args_list):
        """input-gradient chain whose running sum stays in tensor memory (ssr_conv_tc_chain_acc)"""
        arr = (L.ConvTcArgs * len(args_list))(*args_list)
        self.keep.append(arr)
        self._emit(lib().ssr_conv_tc_chain_acc, (arr, len(args_list)))

    def fork(self):
        self.calls.append((_FORK,))

    def join(self):
        self.calls.append((_JOIN,))

    def side(self):
        plan = self

        class _Side:
            def __enter__(self_):
                plan._lane = 1

            def __exit__(self_, *exc):
                plan._lane = 0
        return _Side()

    def extend(self, other):
        self.calls.extend(other.calls)
        self.keep.extend(other.keep)
        self.has_side = self.has_side or other.has_side

    def run(self, stream=None, lane=None):
        s = stream if stream is not None else cur_stream()
        if not self.has_side:
            for fn, args in self.calls:
                rc = fn(*args, s)
                if rc != 0:
                    L.check(rc)
            return
        if lane is None:
            lane = SideLane.get()
        for call in self.calls:
            tag = call[0]
            if tag is _SIDE:
                rc = call[1](*call[2], lane.handle)
            elif tag is _FORK:
                lane.fork(s)
                continue
            elif tag is _JOIN:
                lane.join(s)
                continue
            else:
                rc = tag(*call[1], s)
            if rc != 0:
                L.check(rc)
        lane.join(s)

    def main_calls(self):
        """(fn, args) of the main-lane launches only"""
        return [c for c in self.calls if c[0] not in (_FORK, _JOIN, _SIDE)]

    def __len__(self):
        return sum(1 for c in self.calls if c[0] not in (_FORK, _JOIN))


class PackedConv:
    """Packed bf16 tensor-core operand(s) of one convolution (forward and, optionally, input-gradient form)."""

    def __init__(self, weight, bias, cin_buf, want_dgrad, device, inv_scale=None, cout_buf=None):
        # weight: f32 OIHW tensor view living in the owner's flat parameter buffer
        self.weight, self.bias = weight, bias
        self.cout, self.cin, self.r, _ = weight.shape
        self.cin_buf = cin_buf                      # channels the forward conv reads (multiple of 16)
        self.inv_scale = inv_scale
        n_pad = C.c_int32(0)
        self.k_pad = round_up(cin_buf, 64)
        nbytes = lib().ssr_packed_weight_bytes(self.k_pad, self.cout, self.r, C.byref(n_pad))
        self.n_pad = n_pad.value
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.packed_dg = None
        if want_dgrad:
            # reduce dim = conv output channels as stored in the gradient buffer (multiple of 16)
            self.cout_buf = cout_buf or round_up(self.cout, 16)
            self.k_pad_dg = round_up(self.cout_buf, 64)
            nbytes = lib().ssr_packed_weight_bytes(self.k_pad_dg, self.cin, self.r, C.byref(n_pad))
            self.n_pad_dg = n_pad.value
            self.packed_dg = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def descs(self):
        out = []
        d = PackDesc()
        d.w = self.weight.data_ptr()
        d.dst = self.packed.data_ptr()
        d.inv_scale = self.inv_scale.data_ptr() if self.inv_scale is not None else None
        d.cout, d.cin, d.r, d.mode, d.k_pad, d.n_pad = self.cout, self.cin, self.r, L.PACK_FWD, self.k_pad, self.n_pad
        out.append(d)
        if self.packed_dg is not None:
            d = PackDesc()
            d.w = self.weight.data_ptr()
            d.dst = self.packed_dg.data_ptr()
            # dg_inv_scale: the input-gradient operand may carry a constant factor (1 / dg_inv_scale) of its own
            dg_inv = getattr(self, "dg_inv_scale", None)
            d.inv_scale = dg_inv.data_ptr() if dg_inv is not None else (self.inv_scale.data_ptr() if self.inv_scale is not None else None)
            # a 4 x 4 kernel is the stride-2 conv: its input gradient runs as four 2 x 2 parity-class convs (SSR_PACK_DGRAD_S2)
            d.cout, d.cin, d.r, d.mode, d.k_pad, d.n_pad = (self.cout, self.cin, self.r, L.PACK_DGRAD_S2 if self.r == 4 else L.PACK_DGRAD,
                                                            self.k_pad_dg, self.n_pad_dg)
            out.append(d)
        return out


def dgrad_s2_class_ptr(conv, cls):
    """device pointer of parity class `cls` (= oy * 2 + ox) inside the SSR_PACK_DGRAD_S2 operand of a 4 x 4 stride-2 conv:
    [class][cout chunk][2 x 2 tap][n_pad][64] bf16"""
    per_class = (conv.k_pad_dg // 64) * 4 * conv.n_pad_dg * 64 * 2
    return conv.packed_dg.data_ptr() + cls * per_class


class Packer:
    """All PackedConv of a network -> one ssr_pack_conv_weights_batched launch."""

    def __init__(self, convs, device):
        descs = []
        for cv in convs:
            descs.extend(cv.descs())
        arr = (PackDesc * len(descs))(*descs)
        raw = bytes(arr)
        self.n = len(descs)
        self.has_gemm = int(any(d.mode in (L.PACK_FWD_GEMM, L.PACK_DGRAD_GEMM) for d in descs))
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        # flat work list {descriptor, tile}: one thread block per 8-row x 64-K tile (ssr_pack_conv_weights_tiled)
        self.work = None
        if not self.has_gemm:
            work = []
            for i, d in enumerate(descs):
                work.extend((i, t) for t in range(lib().ssr_pack_tile_count(d.k_pad, d.n_pad)))
            self.n_work = len(work)
            self.work = torch.tensor(work, dtype=torch.int32).reshape(-1, 2).contiguous().to(device)

    def run(self, stream=None):
        s = stream if stream is not None else cur_stream()
        if self.work is not None:
            L.check(lib().ssr_pack_conv_weights_tiled(self.table.data_ptr(), self.work.data_ptr(), self.n_work, s))
        else:
            L.check(lib().ssr_pack_conv_weights_batched(self.table.data_ptr(), self.n, self.has_gemm, s))


def conv_args(x_ptr, B, H, W, x_stride, cin, w_ptr, r, cout, n_pad, bias=None, act=0, s0=1.0,
              res1=None, res1_kind=L.SSR_BF16, res1_stride=0, s1=0.0,
              res2=None, res2_kind=L.SSR_BF16, res2_stride=0, s2=0.0,
              mask=None, mask_stride=0, mask_lo=0, mask_relu=0,
              out=None, out_stride=0, out32=None, out32_mode=L.OUT32_NONE, out32_stride=0,
              n_tile=0, mt=0, splits=0, res1_cmax=0, out_lo=0, bias_grad=None, bias_grad_scale=1.0,
              stride=0, pad_y=0, pad_x=0, out_oy=0, out_ox=0):
    a = L.ConvTcArgs()
    a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cin = x_ptr, B, H, W, x_stride, cin
    a.w_packed, a.r, a.cout, a.n_pad = w_ptr, r, cout, n_pad
    a.bias = bias
    a.act, a.s0 = act, s0
    if res1 is not None:
        a.res1, a.res1_kind, a.res1_pix_stride, a.s1 = res1, res1_kind, res1_stride, s1
    if res2 is not None:
        a.res2, a.res2_kind, a.res2_pix_stride, a.s2 = res2, res2_kind, res2_stride, s2
    if mask is not None:
        a.mask, a.mask_pix_stride, a.mask_lo, a.mask_relu = mask, mask_stride, mask_lo, mask_relu
    if out is not None:
        a.out_bf16, a.out_pix_stride = out, out_stride
    if out32 is not None:
        a.out_f32, a.out32_mode, a.out32_pix_stride = out32, out32_mode, out32_stride
    a.n_tile, a.mt, a.splits, a.res1_cmax = n_tile, mt, splits, res1_cmax
    a.out_lo = out_lo
    if bias_grad is not None:
        a.bias_grad, a.bias_grad_scale = bias_grad, bias_grad_scale
    a.stride, a.pad_y, a.pad_x, a.out_oy, a.out_ox = stride, pad_y, pad_x, out_oy, out_ox
    return a


# --------------------------------------------------------------------------------------------- flat parameter storage
class FlatBuffer:
    """One contiguous f32 device buffer holding every tensor of a network (each aligned to 256 bytes), so that the
    optimiser, the EMA and the gradient all-reduce are single launches over one pointer."""

    ALIGN = 64  # floats

    def __init__(self, shapes, device):
        self.offsets = {}
        off = 0
        for name, shp in shapes.items():
            n = 1
            for s in shp:
                n *= s
            self.offsets[name] = (off, n, tuple(shp))
            off += round_up(n, self.ALIGN)
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)

    def view(self, name):
        off, n, shp = self.offsets[name]
        return self.flat[off:off + n].view(shp)

    def views(self):
        return {k: self.view(k) for k in self.offsets}

    def like(self):
        other = FlatBuffer.__new__(FlatBuffer)
        other.offsets, other.numel = self.offsets, self.numel
        other.flat = torch.zeros_like(self.flat)
        return other


class GemmConv:
    """A k x k strided conv computed as im2col + 1x1 GEMM (the 4x4 stride-2 discriminator convs)."""

    def __init__(self, weight, want_dgrad, device, inv_scale=None):
        self.weight, self.bias = weight, None
        self.cout, self.cin, self.r, _ = weight.shape
        self.inv_scale = inv_scale
        self.kk = self.r * self.r * self.cin
        n_pad = C.c_int32(0)
        self.k_pad = round_up(self.kk, 64)
        nbytes = lib().ssr_packed_weight_bytes(self.k_pad, self.cout, 1, C.byref(n_pad))
        self.n_pad = n_pad.value
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.packed_dg = None
        if want_dgrad:
            self.k_pad_dg = round_up(self.cout, 64)
            nbytes = lib().ssr_packed_weight_bytes(self.k_pad_dg, self.kk, 1, C.byref(n_pad))
            self.n_pad_dg = n_pad.value
            self.packed_dg = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def descs(self):
        out = []
        for mode, dst, k_pad, n_pad in ((L.PACK_FWD_GEMM, self.packed, self.k_pad, self.n_pad),
                                        (L.PACK_DGRAD_GEMM, self.packed_dg, getattr(self, "k_pad_dg", 0),
                                         getattr(self, "n_pad_dg", 0))):
            if dst is None:
                continue
            d = PackDesc()
            d.w, d.dst = self.weight.data_ptr(), dst.data_ptr()
            d.inv_scale = self.inv_scale.data_ptr() if self.inv_scale is not None else None
            d.cout, d.cin, d.r, d.mode, d.k_pad, d.n_pad = self.cout, self.cin, self.r, mode, k_pad, n_pad
            out.append(d)
        return out


class WgradSet:
    """f32 accumulators [taps][cx_rows][cy_stride] for every conv of a network + the batched unpack into OIHW grads."""

    def __init__(self, device):
        self.device = device
        self.items = []   # (name, conv, cx_rows, cy_stride, offset)
        self.total = 0
        self.acc = None
        self.table = None

    def add(self, name, conv, cx_rows):
        cy_stride = round_up(conv.cout, 4)
        n = conv.r * conv.r * cx_rows * cy_stride
        self.items.append((name, conv, cx_rows, cy_stride, self.total))
        self.total += round_up(n, 64)

    def finalize(self, grad_of, accumulate=1):
        from ._protos import UnpackDesc
        self.acc = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.slot = {}
        descs = []
        for name, conv, cx_rows, cy_stride, off in self.items:
            self.slot[name] = (self.acc.data_ptr() + 4 * off, cx_rows, cy_stride)
            d = UnpackDesc()
            d.acc = self.acc.data_ptr() + 4 * off
            d.grad = grad_of(name).data_ptr()
            d.cx_rows, d.acc_stride, d.cout, d.cin, d.r = cx_rows, cy_stride, conv.cout, conv.cin, conv.r
            d.accumulate, d.scale = accumulate, 1.0
            descs.append(d)
        arr = (UnpackDesc * len(descs))(*descs)
        self.n = len(descs)
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)

    def args(self, name, x_ptr, x_stride, cx, dy_ptr, dy_stride, cy, B, H, W, r, scale=1.0):
        from ._protos import WgradArgs
        ptr, cx_rows, cy_stride = self.slot[name]
        a = WgradArgs()
        a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cx = x_ptr, B, H, W, x_stride, cx
        a.dy, a.dy_pix_stride, a.cy, a.r = dy_ptr, dy_stride, cy, r
        # r == 1 on a k x k conv = the im2col GEMM form: rows are (tap, ci) flattened, i.e. the same memory
        a.out, a.out_cx_rows, a.out_stride, a.scale, a.splits = ptr, max(cx_rows, cx), cy_stride, scale, 0
        return a

    def zero(self):
        self.acc.zero_()

    def unpack(self, stream=None):
        L.check(lib().ssr_wgrad_unpack_batched(self.table.data_ptr(), self.n, stream if stream is not None else cur_stream()))


def plan_wgrad(plan, args):
    plan.keep.append(args)
    plan._emit(lib().ssr_wgrad_tc, (C.byref(args),))


def allreduce_sum_(flat, process_group=None):
    """sum all-reduce of one flat gradient buffer (the averaging 1/world is folded into the fused Adam kernel)"""
    import torch.distributed as dist
    if process_group is None and not (dist.is_available() and dist.is_initialized()):
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=process_group)
    return flat


def broadcast_from_rank0_(tensors, process_group=None):
    """in-place broadcast of every tensor from the group's rank 0 (what DistributedDataParallel does to parameters and
    buffers when it wraps a module)"""
    import torch.distributed as dist
    if process_group is None and not (dist.is_available() and dist.is_initialized()):
        return tensors
    src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
    for t in tensors:
        dist.broadcast(t, src=src, group=process_group)
    return tensors


def rank_slice(n_items, rank, world):
    """contiguous shard of `n_items` independent work items (tiles / chunks) for `rank`: no collective needed"""
    per = (n_items + world - 1) // world
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def plan_wgrad_batch(plan, args_list):
    """several independent weight-gradient problems -> one ssr_wgrad_tc_batched call"""
    from ._protos import WgradArgs
    arr = (WgradArgs * len(args_list))(*args_list)
    plan.keep.append(arr)
    plan._emit(lib().ssr_wgrad_tc_batched, (arr, len(args_list)))
