# Hardware probe for the next conv layout (run under gpurun):
#   1. SSR_CONV_TW=8            8-pixel-wide tiles (16 image rows per M tile), contiguous operand rows, SBO = 1024 B  -> must pass
#   2. + SSR_DBG_PITCH=2        the same tiles inside 10-pixel-wide shared-memory rows: the M = 128 operand window is 16 groups
#                               of 8 rows at SBO = 1280 B.  If the conv tests still pass, the 128B swizzle is a pure function of the
#                               row address for ANY group stride, and a tile WITH its halo columns can be loaded (or kept resident)
#                               once and serve all nine taps through descriptor offsets (kx = +128 B, ky = +pitch * 128 B).
set -x
SSR_CONV_TW=8 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -x -q 2>&1 | tail -4
SSR_CONV_TW=8 SSR_DBG_PITCH=2 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -x -q 2>&1 | tail -4
