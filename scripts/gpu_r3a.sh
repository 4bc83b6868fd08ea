# GPU call 3A: narrow forward epilogue (registers pinned, early accumulator release)
set -x
O=gpurun_out/r3a; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -x -q > $O/conv_tests.log 2>&1; tail -n 3 $O/conv_tests.log
timeout 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1; cat $O/bench_conv_big.log
SSR_CONV_DBG=2 timeout 200 python scripts/bench_conv_big.py "conv" 2>&1 | grep -E "G tail|D conv0|VGG conv1_2" > $O/dbg2.log; cat $O/dbg2.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
