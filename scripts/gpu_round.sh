#!/bin/bash
# One GPU-box visit: the changed kernels' tests first (short timeout), the suite (without the 2.5-minute trajectory test) and the bench.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "several") > $O/t_new.log 2>&1; echo "rc=$?" >> $O/t_new.log; tail -4 $O/t_new.log
(time timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py::test_trajectory_20_steps) > $O/t_suite.log 2>&1; echo "rc=$?" >> $O/t_suite.log; tail -6 $O/t_suite.log
(timeout 300 python bench.py --no-cpu-baseline --no-extras) > $O/bench_f.json 2> $O/bench_f.err; cut -c1-160 $O/bench_f.json
(SSR_RDB_FUSE=6 timeout 300 python bench.py --no-cpu-baseline --no-extras) > $O/bench_f6.json 2> $O/bench_f6.err; cut -c1-160 $O/bench_f6.json
