#!/bin/bash
# One GPU-box visit: the new kernels' tests first (short timeouts: a deadlock must not eat the budget), then the suite, the bench
# (with and without multi-block resident launches) and the stock-library baseline.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "several or chain") > $O/t_new.log 2>&1; echo "rc=$?" >> $O/t_new.log; tail -4 $O/t_new.log
(time timeout 300 python -m pytest tests/test_generator_gpu.py -x -q -s -k "split_bf16") > $O/t_tight.log 2>&1; echo "rc=$?" >> $O/t_tight.log; grep -E "blocks=|passed|failed|rc=" $O/t_tight.log | tail -8
(time timeout 200 python -m pytest tests/test_variants_gpu.py -x -q -k "ssim") > $O/t_ssim.log 2>&1; echo "rc=$?" >> $O/t_ssim.log; tail -4 $O/t_ssim.log
(time timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_depth_parity_gpu.py::test_trajectory_20_steps) > $O/t_suite.log 2>&1; echo "rc=$?" >> $O/t_suite.log; tail -4 $O/t_suite.log
(timeout 300 python bench.py --no-cpu-baseline) > $O/bench_fuse3.json 2> $O/bench_fuse3.err; tail -c 300 $O/bench_fuse3.json
(SSR_RDB_FUSE=1 timeout 300 python bench.py --no-cpu-baseline) > $O/bench_fuse1.json 2> $O/bench_fuse1.err; tail -c 300 $O/bench_fuse1.json
(timeout 400 python scripts/library_baseline.py --out $O/library_baseline.json) > $O/library_baseline.log 2>&1; tail -8 $O/library_baseline.log
