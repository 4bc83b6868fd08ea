#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest tests/test_modules_gpu.py -q -s -k cuda_graph_replay 2>&1 | grep -E "step-2 grad|passed|failed" ; done > $O/t_graph.log 2>&1; cat $O/t_graph.log
(time timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py::test_trajectory_20_steps) > $O/t_suite.log 2>&1; echo "rc=$?" >> $O/t_suite.log; tail -6 $O/t_suite.log
