set -x
O=gpurun_out/r2j; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q > $O/conv_tests.log 2>&1; echo "rc=$?" >> $O/conv_tests.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py --deselect tests/test_conv_tc_gpu.py > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
SSR_CONV_WSTAT=0 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nows.json 2> $O/bench_nows.err
timeout -s KILL 300 python scripts/bench_conv.py 32 > $O/bench_conv.log 2>&1
timeout -s KILL 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1
tail -3 $O/conv_tests.log $O/tests.log; cut -c1-200 $O/bench.json $O/bench_nows.json; grep -E "hr 128|up1|conv0|conv1_2|conv6" $O/bench_conv.log $O/bench_conv_big.log
