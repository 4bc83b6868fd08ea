# GPU call N: per-shape A/B of the halo form
set -x
O=gpurun_out/r2n; mkdir -p $O
timeout 300 python scripts/bench_conv_big.py > $O/halo1.log 2>&1
SSR_CONV_HALO=0 timeout 300 python scripts/bench_conv_big.py > $O/halo0.log 2>&1
SSR_CONV_HALO=0 SSR_CONV_WSTAT=0 timeout 300 python scripts/bench_conv_big.py > $O/halo0_ws0.log 2>&1
SSR_CONV_MT=1 timeout 300 python scripts/bench_conv_big.py > $O/halo1_mt1.log 2>&1
paste -d'|' $O/halo1.log $O/halo0.log | cut -c1-90,150-190
