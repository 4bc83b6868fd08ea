# GPU call Z: decomposition of the epilogue-alone time (no MMAs): 2 = all, 6 = no stores, 10 = no TMEM loads, 14 = arithmetic only
set -x
O=gpurun_out/r2z; mkdir -p $O
for d in 2 6 10 14; do
  echo "== SSR_CONV_DBG=$d"
  SSR_CONV_DBG=$d timeout 200 python scripts/bench_conv_big.py "conv" 2>&1 | grep -E "G tail|D conv6 dgrad|VGG conv2_1 |D conv0"
done > $O/dbg.log 2>&1
cat $O/dbg.log
