set -x
O=gpurun_out/r2f; mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_wgrad_tc_gpu.py -q > $O/wgrad_fuse3.log 2>&1; echo "rc=$?" >> $O/wgrad_fuse3.log
SSR_WGRAD9_FUSE3=0 timeout -s KILL 300 python -m pytest tests/test_wgrad_tc_gpu.py -q > $O/wgrad_plain.log 2>&1; echo "rc=$?" >> $O/wgrad_plain.log
timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py -q -x -k "chain" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py 2>&1 | grep -v "^    layer\|^  CTA" > $O/timeline.log; echo "rc=$?" >> $O/timeline.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
SSR_WGRAD9_FUSE3=0 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nofuse.json 2> $O/bench_nofuse.err
timeout -s KILL 120 python scripts/bench_rdb_bwd.py > $O/rdb_bwd.log 2>&1
tail -4 $O/*.log; cut -c1-300 $O/bench.json $O/bench_nofuse.json
