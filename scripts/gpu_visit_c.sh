#!/bin/bash
# Round-2c last visit: side-lane tests (incl. the per-phase-graph capture path world > 1 uses), graph-vs-eager, and the default bench line of the shipped tree
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
(time timeout 100 python -m pytest tests/test_overlap_gpu.py tests/test_modules_gpu.py -x -q -s -k "side_lane or cuda_graph_replay") > $O/t_overlap_c.log 2>&1; echo "rc=$?" >> $O/t_overlap_c.log; tail -14 $O/t_overlap_c.log
(time timeout 100 python bench.py --no-cpu-baseline --no-extras) > $O/bench_shipped_n1.json 2> $O/bench_shipped_n1.err; cut -c1-200 $O/bench_shipped_n1.json
