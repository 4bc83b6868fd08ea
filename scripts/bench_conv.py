"""Per-shape device time of ssr_conv_tc / ssr_wgrad_tc (50 back-to-back launches captured in one CUDA graph, warm L2)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import _lib as L
from satlas_super_resolution_b200._protos import WgradArgs
from satlas_super_resolution_b200.ops import conv_args, cur_stream

lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""   # substring filter on the case name (for ncu captures)


def time_graph(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def conv_case(cin, cout, H, W, mt=0, stride=192):
    x = torch.randn(B, H, W, stride, device="cuda").to(torch.bfloat16)
    n_pad = C.c_int32(0)
    kp = (cin + 63) // 64 * 64
    nbytes = lib.ssr_packed_weight_bytes(kp, cout, 3, C.byref(n_pad))
    wp = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, H, W, max(cout, 64), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(max(cout, 64), device="cuda")
    a = conv_args(x.data_ptr(), B, H, W, stride, cin, wp.data_ptr(), 3, cout, n_pad.value, bias=bias.data_ptr(), act=1,
                  out=out.data_ptr(), out_stride=max(cout, 64), mt=mt)
    us = time_graph(lambda: L.check(lib.ssr_conv_tc(C.byref(a), cur_stream())))
    fl = 2.0 * B * H * W * cout * 9 * cin
    return us, fl / us / 1e6


def wgrad_case(cx, cy, H, W):
    x = torch.randn(B, H, W, 192, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H, W, 64, device="cuda").to(torch.bfloat16)
    acc = torch.zeros(9 * cx * 64, device="cuda")
    a = WgradArgs()
    a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cx = x.data_ptr(), B, H, W, 192, cx
    a.dy, a.dy_pix_stride, a.cy, a.r = dy.data_ptr(), 64, cy, 3
    a.out, a.out_cx_rows, a.out_stride, a.scale, a.splits = acc.data_ptr(), cx, 64, 1.0, 0
    us = time_graph(lambda: L.check(lib.ssr_wgrad_tc(C.byref(a), cur_stream())))
    fl = 2.0 * B * H * W * cy * 9 * cx
    return us, fl / us / 1e6


print(f"B={B}")
for name, cin, cout, hw in (("rdb conv1", 64, 32, 32), ("rdb conv2", 96, 32, 32), ("rdb conv3", 128, 32, 32), ("rdb conv4", 160, 32, 32),
                            ("rdb conv5", 192, 64, 32), ("dgrad5", 64, 192, 32), ("dgrad4", 32, 160, 32), ("dgrad3", 32, 128, 32),
                            ("dgrad2", 32, 96, 32), ("dgrad1", 32, 64, 32), ("hr 128^2", 64, 64, 128), ("up1 64^2", 64, 64, 64)):
    if ONLY not in name:
        continue
    for mt in (1, 2):
        us, tf = conv_case(cin, cout, hw, hw, mt)
        print(f"conv  {name:10s} {cin:3d}->{cout:3d} {hw:3d}^2 mt={mt}: {us:7.1f} us  {tf:7.1f} TFLOP/s")
for name, cx, cy, hw in (("rdb conv1", 64, 32, 32), ("rdb conv3", 128, 32, 32), ("rdb conv4", 160, 32, 32), ("rdb conv5", 192, 64, 32),
                         ("hr 128^2", 64, 64, 128)):
    if ONLY not in name:
        continue
    us, tf = wgrad_case(cx, cy, hw, hw)
    print(f"wgrad {name:10s} {cx:3d}x{cy:3d} {hw:3d}^2: {us:7.1f} us  {tf:7.1f} TFLOP/s")
