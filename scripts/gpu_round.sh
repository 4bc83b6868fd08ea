#!/bin/bash
# One GPU-box visit: the suite (without the 2.5-minute trajectory test), the bench and the launch list of one step.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py::test_trajectory_20_steps) > $O/t_suite.log 2>&1; echo "rc=$?" >> $O/t_suite.log; tail -6 $O/t_suite.log
(timeout 300 python bench.py --no-cpu-baseline --no-extras) > $O/bench_e.json 2> $O/bench_e.err; cut -c1-160 $O/bench_e.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram_e.csv python scripts/profile_step.py 32 > $O/ncu_list_e.log 2>&1
