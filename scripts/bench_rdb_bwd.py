"""Times the batched weight-gradient launch of one dense block (B=32, 32x32) -- run once per SSR_WGRAD_BATCH_CTAS value."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from satlas_super_resolution_b200 import _lib as L  # noqa: E402
from satlas_super_resolution_b200._protos import WgradArgs  # noqa: E402

lib = L.load()
B, H, W, nf, g = 32, 32, 32, 64, 32
cw = nf + 4 * g
cur = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)
dg = torch.randn(B, H, W, cw).cuda().to(torch.bfloat16)
xin = torch.randn(B, H, W, nf).cuda().to(torch.bfloat16)
arr = (WgradArgs * 5)()
keep = []
for i, k in enumerate(range(5, 0, -1)):
    cx = nf + (k - 1) * g
    cy = nf if k == 5 else g
    acc = torch.zeros(9 * cx * cy, device="cuda")
    keep.append(acc)
    a = arr[i]
    a.x, a.n_img, a.h, a.w, a.x_pix_stride, a.cx = cur.data_ptr(), B, H, W, cw, cx
    if k == 5:
        a.dy, a.dy_pix_stride = xin.data_ptr(), nf
    else:
        a.dy, a.dy_pix_stride = dg.data_ptr() + 2 * cx, cw
    a.cy, a.r, a.out, a.out_cx_rows, a.out_stride, a.scale, a.splits = cy, 3, acc.data_ptr(), cx, cy, 1.0, 0
s = torch.cuda.current_stream().cuda_stream


def run():
    L.check(lib.ssr_wgrad_tc_batched(arr, 5, s))


for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
print(f"SSR_WGRAD_BATCH_CTAS={os.environ.get('SSR_WGRAD_BATCH_CTAS', 'default')}: batched wgrad of one dense block {e0.elapsed_time(e1) / n * 1e3:.1f} us")
