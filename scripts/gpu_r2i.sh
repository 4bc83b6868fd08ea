set -x
O=gpurun_out/r2i; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py 2>&1 | grep -v "^    layer\|^  CTA" > $O/timeline.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
timeout -s KILL 300 python scripts/bench_conv.py 32 > $O/bench_conv.log 2>&1
timeout -s KILL 300 python scripts/bench_conv_big.py > $O/bench_conv_big.log 2>&1
timeout -s KILL 300 python -m pytest tests/test_infer_gpu.py tests/test_generator_gpu.py -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_infer.json 2> $O/bench_infer.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram.csv python scripts/profile_step.py 32 > $O/ncu_list.log 2>&1
tail -3 $O/chain_tests.log; cut -c1-200 $O/bench.json; cat $O/bench_conv.log $O/bench_conv_big.log; tail -3 $O/tests.log; cut -c1-300 $O/bench_infer.json
