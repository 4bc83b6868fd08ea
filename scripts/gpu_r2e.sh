set -x
O=gpurun_out/r2e; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline.log 2>&1; echo "rc=$?" >> $O/timeline.log
SSR_RDB_MULTICAST=1 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain_equals_plain or chain_acc" > $O/chain_tests_mc.log 2>&1; echo "rc=$?" >> $O/chain_tests_mc.log
SSR_RDB_MULTICAST=1 SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline_mc.log 2>&1; echo "rc=$?" >> $O/timeline_mc.log
timeout -s KILL 600 python -m pytest tests/test_train_gpu.py tests/test_generator_gpu.py tests/test_variants_gpu.py tests/test_metrics_gpu.py -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
SSR_RDB_MULTICAST=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_mc.json 2> $O/bench_mc.err
tail -4 $O/*.log; cut -c1-300 $O/bench.json $O/bench_mc.json
