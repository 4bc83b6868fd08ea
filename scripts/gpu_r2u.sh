# GPU call U: resident kernel with division-free ring counters and immediate weight-tap offsets
set -x
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_generator_gpu.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline.log 2>&1; echo "rc=$?" >> $O/timeline.log
grep -E "^==|first stamp" $O/timeline.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench2.json 2> $O/bench2.err; cut -c1-200 $O/bench2.json
