"""Hardware probe: does a UMMA smem descriptor whose start address is offset by a multiple of 128 B that is NOT a
multiple of the 1024-B swizzle atom still address the rows TMA wrote (i.e. is the 128B swizzle a function of the
absolute shared-memory address)?  If yes, all nine 3x3 taps can be served from ONE halo tile in shared memory."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch.nn.functional as F

for off in (0, 1, 3, 8):
    os.environ["SSR_DBG_AOFF"] = str(off)
    from test_conv_tc_gpu import run_conv, make
    x, w = make(1, 64, 32, 1, 256, r=1, seed=1)
    got, _ = run_conv(x, w, out_kind="f32")          # [1,32,1,256]
    ref = F.conv2d(x, w)
    ok_rows, bad_rows = 0, 0
    for tile in range(2):
        for m in range(128 - off):
            g = got[0, :, 0, tile * 128 + m]
            r = ref[0, :, 0, tile * 128 + m + off]
            if torch.allclose(g, r, atol=1e-3, rtol=1e-3): ok_rows += 1
            else: bad_rows += 1
    print(f"SSR_DBG_AOFF={off}: rows matching shifted reference: {ok_rows}, not matching: {bad_rows}")
os.environ["SSR_DBG_AOFF"] = "0"
