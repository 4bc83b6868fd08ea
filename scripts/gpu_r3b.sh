# GPU call 3B: 256-bit epilogue stores (full sectors)
set -x
O=gpurun_out/r3b; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q > $O/conv_tests.log 2>&1; tail -n 3 $O/conv_tests.log
timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1; cat $O/bench_conv_big.log
SSR_CONV_ST256=0 timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big_st128.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
SSR_CONV_ST256=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_st128.json 2> $O/bench_st128.err; cut -c1-200 $O/bench_st128.json
