set -x
O=gpurun_out/r2g; mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -q -k "strided" > $O/strided_conv.log 2>&1; echo "rc=$?" >> $O/strided_conv.log
timeout -s KILL 300 python -m pytest tests/test_wgrad_tc_gpu.py -q > $O/wgrad.log 2>&1; echo "rc=$?" >> $O/wgrad.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -4 $O/*.log; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
