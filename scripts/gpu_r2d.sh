set -x
O=gpurun_out/r2d; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline.log 2>&1; echo "rc=$?" >> $O/timeline.log
timeout -s KILL 300 python -m pytest tests/test_metrics_gpu.py tests/test_generator_gpu.py -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -4 $O/*.log
