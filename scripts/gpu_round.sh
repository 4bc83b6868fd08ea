#!/bin/bash
# One GPU-box visit: the changed kernels' tests first (short timeouts: a deadlock must not eat the budget), then the suite and the
# bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 300 python -m pytest tests/test_resample_gpu.py tests/test_infer_gpu.py -x -q) > $O/t_new.log 2>&1; echo "rc=$?" >> $O/t_new.log; tail -4 $O/t_new.log
(time timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_depth_parity_gpu.py::test_trajectory_20_steps) > $O/t_suite.log 2>&1; echo "rc=$?" >> $O/t_suite.log; tail -4 $O/t_suite.log
(timeout 300 python bench.py --no-cpu-baseline) > $O/bench_c.json 2> $O/bench_c.err; tail -c 300 $O/bench_c.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram_c.csv python scripts/profile_step.py 32 > $O/ncu_list_c.log 2>&1
