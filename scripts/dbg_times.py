import ctypes as C, sys, os
import torch
sys.path.insert(0, ".")
from satlas_super_resolution_b200 import _lib as L
from satlas_super_resolution_b200.ops import conv_args, cur_stream
lib = L.load()
B = 32
for cin, cout, hw in ((64, 32, 32), (192, 64, 32), (64, 192, 32), (64, 64, 128)):
    x = torch.randn(B, hw, hw, 192, device="cuda").to(torch.bfloat16)
    n_pad = C.c_int32(0)
    nbytes = lib.ssr_packed_weight_bytes((cin + 63) // 64 * 64, cout, 3, C.byref(n_pad))
    wp = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, hw, hw, max(cout, 64), device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(256, device="cuda")
    a = conv_args(x.data_ptr(), B, hw, hw, 192, cin, wp.data_ptr(), 3, cout, n_pad.value, bias=bias.data_ptr(), act=1,
                  out=out.data_ptr(), out_stride=max(cout, 64))
    print(f"--- {cin}->{cout} {hw}^2", flush=True)
    for _ in range(4):
        L.check(lib.ssr_conv_tc(C.byref(a), cur_stream()))
    torch.cuda.synchronize()
