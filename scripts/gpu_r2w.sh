# GPU call W: final round-2 profile (launch list + DRAM bytes, ncu --set full of the dominant kernels) and the default bench line
set -x
O=gpurun_out/r2prof; mkdir -p $O
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-300 $O/bench_default.json
bash scripts/profile_round2.sh > $O/profile.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o $O/r02_conv64_halo python scripts/bench_conv_big.py "VGG conv1_2" > $O/ncu_conv64_halo.log 2>&1
timeout 300 python scripts/bench_conv.py 32 > $O/bench_conv.log 2>&1
timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline.log 2>&1
ls -la $O
