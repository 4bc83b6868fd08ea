# GPU call M: halo-tile / stationary-weight 3x3 convs
set -x
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q > $O/conv_tests.log 2>&1; tail -n 5 $O/conv_tests.log
timeout 300 python scripts/bench_conv.py 32 > $O/bench_conv.log 2>&1
timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1
grep -E "hr 128|up1" $O/bench_conv.log; cat $O/bench_conv_big.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_depth_parity_gpu.py --deselect tests/test_conv_tc_gpu.py > $O/tests.log 2>&1; tail -n 3 $O/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
SSR_CONV_HALO=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nohalo.json 2> $O/bench_nohalo.err; cut -c1-200 $O/bench_nohalo.json
