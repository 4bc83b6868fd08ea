# round-2 GPU call B: the new resident dense-block kernel (forward + input-gradient), timelines, full suite, bench
set -x
O=gpurun_out/r2b; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/timeline.log 2>&1; echo "rc=$?" >> $O/timeline.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
SSR_CONV_RESIDENT=0 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nores.json 2> $O/bench_nores.err
tail -4 $O/*.log; cut -c1-400 $O/bench.json; tail -5 $O/bench.err
