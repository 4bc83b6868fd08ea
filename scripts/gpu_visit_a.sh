#!/bin/bash
# Round-2c visit A: the side-lane (SSR_OVERLAP) tests, its A/B at the benchmarked configuration, and one ncu --set full pass over the
# twelve resident dense-block launches of a step (six forward, six input-gradient: up to twelve blocks each).
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
(time timeout 240 python -m pytest tests/test_overlap_gpu.py tests/test_host_cpu.py -x -q -s -k "side_lane or plan_lanes") > $O/t_overlap.log 2>&1; echo "rc=$?" >> $O/t_overlap.log; tail -12 $O/t_overlap.log
(time timeout 240 python scripts/overlap_check.py 32 23) > $O/overlap_check.json 2> $O/overlap_check.err; echo "rc=$?" >> $O/overlap_check.err; cat $O/overlap_check.json; tail -5 $O/overlap_check.err
(time timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rdb_resident_kernel -c 12 -f -o $O/r02c_rdb_all python scripts/profile_step.py 32) > $O/ncu_full_rdb.log 2>&1; tail -3 $O/ncu_full_rdb.log
ls -la $O
