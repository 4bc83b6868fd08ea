"""Static race check of the generator's backward plan with the side lane (generator._build_backward).

The plan is built on CPU tensors (argument structs only -- nothing is launched) and walked in recorded order: a side-lane launch is
concurrent with every main-lane launch recorded between its fork and the first join after it.  For each such pair the buffers one
writes must not be touched by the other.  This proves the sizing of the rotating buffers (two dY sets, 2 * fuse + 1 block-gradient
and 2 * ceil(fuse / 3) + 3 RRDB-gradient buffers) for aligned and unaligned group sizes, without a GPU."""
import bisect
import ctypes as C
import os
import subprocess
import sys
from collections import OrderedDict

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# --------------------------------------------------------------------------------------------- the checker
def _tensors(obj, seen, out):
    """every torch tensor reachable from obj (attributes, lists, tuples, dicts; ops.Act holds its tensor in .t)"""
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        out.append(obj)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _tensors(o, seen, out)
    elif isinstance(obj, dict):
        for o in obj.values():
            _tensors(o, seen, out)
    elif hasattr(obj, "__dict__") and not isinstance(obj, (type, C.Structure, C.Array)) and not callable(obj):
        for o in vars(obj).values():
            _tensors(o, seen, out)


class Allocations:
    def __init__(self, *roots):
        ts = []
        seen = set()
        for r in roots:
            _tensors(r, seen, ts)
        spans = {}
        for t in ts:
            st = t.untyped_storage()
            if st.nbytes():
                spans[st.data_ptr()] = max(spans.get(st.data_ptr(), 0), st.nbytes())
        self.lo = sorted(spans)
        self.hi = [lo + spans[lo] for lo in self.lo]

    def whole(self, p):
        """(lo, hi) of the allocation holding address p"""
        i = bisect.bisect_right(self.lo, p) - 1
        if i < 0 or p >= self.hi[i]:
            raise AssertionError(f"pointer {p:#x} is in no known allocation: the checker does not see every buffer")
        return (self.lo[i], self.hi[i])

    def maybe(self, p):
        i = bisect.bisect_right(self.lo, p) - 1
        return (self.lo[i], self.hi[i]) if i >= 0 and p < self.hi[i] else None


def accesses(fn, args, lb, alloc):
    """(reads, writes) of one recorded C-ABI call as address ranges: whole allocations, except weight-gradient accumulators (exact slot)"""
    from satlas_super_resolution_b200 import _lib as L
    reads, writes = [], []
    R = lambda p: reads.append(alloc.whole(p)) if p else None
    W = lambda p: writes.append(alloc.whole(p)) if p else None
    structs = lambda a: [a[0]._obj] if len(a) == 1 else list(a[0])[:a[1]]
    if fn in (lb.ssr_conv_tc, lb.ssr_conv_tc_chain, lb.ssr_conv_tc_chain_acc):
        for a in structs(args):
            for p in (a.x, a.w_packed, a.bias, a.res1, a.res2, a.mask):
                R(p)
            for p in (a.out_bf16, a.out_f32, a.bias_grad):
                W(p)
            if a.out_f32 and a.out32_mode == L.OUT32_PLANAR4_ACC:
                R(a.out_f32)
    elif fn in (lb.ssr_wgrad_tc, lb.ssr_wgrad_tc_batched):
        for a in structs(args):
            R(a.x)
            R(a.dy)
            alloc.whole(a.out)
            writes.append((a.out, a.out + a.r * a.r * a.out_cx_rows * a.out_stride * 4))
    elif fn is lb.ssr_bias_grad:
        R(args[0])
        W(args[4])
    else:   # any other kernel: every argument that is an address inside a known buffer counts as read AND written
        for v in args:
            if isinstance(v, int) and v > (1 << 20):
                span = alloc.maybe(v)
                if span:
                    reads.append(span)
                    writes.append(span)
    return reads, writes


def _overlap(xs, ys):
    return any(a[0] < b[1] and b[0] < a[1] for a in xs for b in ys)


def check_plan(plan, lb, alloc):
    """-> (number of concurrent main/side pairs examined, list of conflicts)"""
    from satlas_super_resolution_b200.ops import _FORK, _JOIN, _SIDE
    pending, main_since_fork, forked = [], [], False
    pairs, conflicts = 0, []

    def clash(i_side, s_acc, i_main, m_acc):
        nonlocal pairs
        pairs += 1
        if _overlap(s_acc[1], m_acc[0] + m_acc[1]) or _overlap(s_acc[0], m_acc[1]):
            conflicts.append((i_side, i_main))

    for idx, call in enumerate(plan.calls):
        tag = call[0]
        if tag is _FORK:
            forked, main_since_fork = True, []
        elif tag is _JOIN:
            pending = []
        elif tag is _SIDE:
            assert forked, "a side-lane call before any fork would not be ordered behind anything"
            acc = accesses(call[1], call[2], lb, alloc)
            for j, macc in main_since_fork:
                clash(idx, acc, j, macc)
            pending.append((idx, acc))
        else:
            acc = accesses(tag, call[1], lb, alloc)
            for j, sacc in pending:
                clash(j, sacc, idx, acc)
            if forked:
                main_since_fork.append((idx, acc))
    return pairs, conflicts


def build_backward_plan(num_block, B=2, overlap=True):
    from satlas_super_resolution_b200 import weights
    from satlas_super_resolution_b200.generator import RRDBNetEngine
    from satlas_super_resolution_b200.ops import FlatBuffer, lib
    state = weights.rrdbnet_state(24, 3, num_block=num_block, seed=0)
    buf = FlatBuffer(OrderedDict((k, tuple(v.shape)) for k, v in state.items()), "cpu")
    grads = buf.like()
    eng = RRDBNetEngine(buf.views(), 24, 3, num_block=num_block, want_grad=True, grads=grads.views(), overlap=overlap)
    ws = eng.workspace(B, 32, 32, True)
    ws.bwd = ws._build_backward(eng)
    return eng, ws, Allocations(eng, ws, buf, grads), lib()


# --------------------------------------------------------------------------------------------- tests
def test_checker_flags_a_rewritten_buffer():
    """main rewrites what a side launch still reads -> flagged; with a join in between -> clean"""
    sys.path.insert(0, ROOT)
    from satlas_super_resolution_b200._protos import WgradArgs
    from satlas_super_resolution_b200.ops import Plan, conv_args, lib, plan_wgrad
    lb = lib()
    x, dy, acc, w, dy2 = (torch.zeros(4096, dtype=torch.uint8) for _ in range(5))
    alloc = Allocations([x, dy, acc, w, dy2])

    def wg(dy_t):
        a = WgradArgs()
        a.x, a.dy, a.out, a.r, a.out_cx_rows, a.out_stride = x.data_ptr(), dy_t.data_ptr(), acc.data_ptr(), 3, 4, 4
        return a

    for with_join, other_buffer, expect in ((False, False, True), (True, False, False), (False, True, False)):
        plan = Plan()
        plan.conv(conv_args(x.data_ptr(), 1, 8, 8, 16, 16, w.data_ptr(), 3, 16, 16, out=dy.data_ptr(), out_stride=16))
        plan.fork()
        with plan.side():
            plan_wgrad(plan, wg(dy))
        if with_join:
            plan.join()
        target = dy2 if other_buffer else dy
        plan.conv(conv_args(x.data_ptr(), 1, 8, 8, 16, 16, w.data_ptr(), 3, 16, 16, out=target.data_ptr(), out_stride=16))
        pairs, conflicts = check_plan(plan, lb, alloc)
        assert bool(conflicts) == expect, (with_join, other_buffer, conflicts)
        assert pairs == (0 if with_join else 1)


@pytest.mark.parametrize("num_block", [23, 9, 5])
def test_backward_plan_with_side_lane_has_no_hazard(num_block):
    sys.path.insert(0, ROOT)
    from satlas_super_resolution_b200.ops import lib
    fuse = lib().ssr_rdb_resident_max_blocks(2, 32, 32)
    eng, ws, alloc, lb = build_backward_plan(num_block)
    n_groups = -(-3 * num_block // fuse)
    assert ws.overlap_bwd == (n_groups > 1) and ws.bwd.has_side == (n_groups > 1)
    pairs, conflicts = check_plan(ws.bwd, lb, alloc)
    assert not conflicts, conflicts[:5]
    if n_groups > 1:
        # every group's weight-gradient launches (one per block) but the last group's are concurrent with the next input-gradient launch;
        # the last group's with the conv_first tail (axpby + its weight gradient)
        assert pairs >= 3 * num_block
    # the single-stream plan: nothing concurrent
    _, ws0, alloc0, _ = build_backward_plan(num_block, overlap=False)
    assert not ws0.bwd.has_side and check_plan(ws0.bwd, lb, alloc0) == (0, [])


def test_a_too_small_rotation_would_be_flagged():
    """the same plan with ONE dY set (group parity ignored) must be reported: the checker sees the hazard the second set removes"""
    sys.path.insert(0, ROOT)
    from satlas_super_resolution_b200 import generator
    from satlas_super_resolution_b200.ops import lib
    fuse = lib().ssr_rdb_resident_max_blocks(2, 32, 32)
    if 27 <= fuse:
        pytest.skip("one group only")
    eng, ws, alloc, lb = build_backward_plan(9)
    from satlas_super_resolution_b200.ops import _SIDE
    # alias the second dY set onto the first by rewriting the recorded pointers of the side-lane weight gradients of group 1
    dgs = ws._bwd_keep[5]
    first, second = dgs[:fuse], dgs[fuse:]
    remap = {b.t.data_ptr(): a.t.data_ptr() for a, b in zip(first, second)}
    size = first[0].t.numel() * 2
    n = 0
    for call in ws.bwd.calls:
        structs = []
        if call[0] is _SIDE and call[1] is lb.ssr_wgrad_tc_batched:
            structs = list(call[2][0])[:call[2][1]]
        elif call[0] is lb.ssr_conv_tc_chain_acc:
            structs = list(call[1][0])[:call[1][1]]
        for a in structs:
            for field in ("dy", "x", "out_bf16"):
                p = getattr(a, field, None)
                if p:
                    for base, to in remap.items():
                        if base <= p < base + size:
                            setattr(a, field, to + (p - base))
                            n += 1
    assert n > 0
    _, conflicts = check_plan(ws.bwd, lb, alloc)
    assert conflicts, "aliasing the two dY sets must produce a hazard"


@pytest.mark.parametrize("fuse", [2, 3, 5, 7])
def test_unaligned_group_sizes_in_a_subprocess(fuse):
    """SSR_RDB_FUSE is read once per process: other group sizes (groups that straddle RRDB boundaries) run in a child process"""
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
            "import test_lane_hazards_cpu as t\n"
            "from satlas_super_resolution_b200.ops import lib\n"
            f"assert lib().ssr_rdb_resident_max_blocks(2, 32, 32) == {fuse}\n"
            "for nb in (4, 7):\n"
            "    eng, ws, alloc, lb = t.build_backward_plan(nb)\n"
            "    assert ws.overlap_bwd and ws.bwd.has_side\n"
            "    pairs, conflicts = t.check_plan(ws.bwd, lb, alloc)\n"
            "    assert pairs > 0 and not conflicts, (nb, conflicts[:5])\n"
            "print('hazard-free')\n")
    env = dict(os.environ, SSR_RDB_FUSE=str(fuse))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and "hazard-free" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
