# GPU call Y: does the epilogue overlap the MMAs?  64 -> 64 / 64 -> 128 halo convs with (1) no epilogue, (2) no MMAs, (4) no stores
set -x
O=gpurun_out/r2y; mkdir -p $O
for d in 0 1 2 4 3; do
  echo "== SSR_CONV_DBG=$d"
  SSR_CONV_DBG=$d timeout 200 python scripts/bench_conv_big.py "conv" 2>&1 | grep -E "G tail|D conv6 |D conv6 dgrad|VGG conv1_2|VGG conv2_1 |D conv0"
done > $O/dbg.log 2>&1
cat $O/dbg.log
