# GPU call P: lean epilogue v2 (no divisions / copies), halo form for short inputs
set -x
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q > $O/conv_tests.log 2>&1; tail -n 5 $O/conv_tests.log
timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1; cat $O/bench_conv_big.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_depth_parity_gpu.py --deselect tests/test_conv_tc_gpu.py > $O/tests.log 2>&1; tail -n 3 $O/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o $O/r02_conv64_halo python scripts/bench_conv_big.py "VGG conv1_2" > $O/ncu_conv64_halo.log 2>&1
