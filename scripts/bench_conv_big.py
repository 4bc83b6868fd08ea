"""Per-shape device time of the discriminator / VGG sized convs (ssr_conv_tc, warm L2, 20 launches per graph)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import _lib as L
from satlas_super_resolution_b200.ops import conv_args, cur_stream
lib = L.load()


def time_graph(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""   # substring filter on the case name (for ncu captures)


def case(name, B, cin, cout, H, W, r=3):
    if ONLY not in name:
        return
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    n_pad = C.c_int32(0)
    nbytes = lib.ssr_packed_weight_bytes((cin + 63) // 64 * 64, cout, r, C.byref(n_pad))
    wp = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    a = conv_args(x.data_ptr(), B, H, W, cin, cin, wp.data_ptr(), r, cout, n_pad.value, act=1, out=out.data_ptr(), out_stride=cout)
    us = time_graph(lambda: L.check(lib.ssr_conv_tc(C.byref(a), cur_stream())))
    fl = 2.0 * B * H * W * cout * r * r * cin
    print(f"{name:28s} B={B:3d} {cin:4d}->{cout:4d} {H:4d}x{W:<6d} r={r}: {us:8.1f} us {fl / us / 1e6:7.1f} TFLOP/s")


case("D conv4", 32, 512, 256, 32, 32)
case("D conv5", 32, 256, 128, 64, 64)
case("D conv6", 32, 128, 64, 128, 128)
case("D conv0 (32 in)", 32, 32, 64, 128, 128)
case("D conv4 dgrad", 32, 256, 512, 32, 32)
case("D conv1 gemm", 1, 1024, 128, 1, 131072, r=1)
case("D conv2 gemm", 1, 2048, 256, 1, 32768, r=1)
case("D conv3 gemm", 1, 4096, 512, 1, 8192, r=1)
case("D conv1 dcol gemm", 1, 128, 1024, 1, 131072, r=1)
case("G tail / D conv7-8", 32, 64, 64, 128, 128)
case("G conv_up1", 32, 64, 64, 64, 64)
case("D conv6 dgrad", 32, 64, 128, 128, 128)
case("VGG conv2_1", 64, 64, 128, 64, 64)
case("VGG conv2_1 dgrad", 32, 128, 64, 64, 64)
case("VGG conv1_2", 64, 64, 64, 128, 128)
case("VGG conv2_2", 64, 128, 128, 64, 64)
case("VGG conv3_x", 64, 256, 256, 32, 32)
case("VGG conv4_x", 64, 512, 512, 16, 16)
case("VGG conv5_x", 64, 512, 512, 8, 8)
