#!/bin/bash
# Round-2c visit D: the VGG pass split by depth over the side lane -- tests, then the A/B of the split depths at the benchmarked configuration
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
(time timeout 90 python -m pytest tests/test_overlap_gpu.py tests/test_train_gpu.py -x -q -s -k "side_lane or perceptual") > $O/t_overlap_d.log 2>&1; echo "rc=$?" >> $O/t_overlap_d.log; tail -16 $O/t_overlap_d.log
(time timeout 100 python scripts/overlap_check.py 32 23 1:all/1/1:pool2/1:pool4 timing) > $O/overlap_variants_d.json 2> $O/overlap_variants_d.err; echo "rc=$?" >> $O/overlap_variants_d.err; cat $O/overlap_variants_d.json; tail -3 $O/overlap_variants_d.err
