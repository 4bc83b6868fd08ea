#!/bin/bash
# Final visit of round 2c: the whole GPU suite exactly as the driver runs it, smoke(), the default bench line + the reference arm,
# the side-lane A/B at the benchmarked configuration, and the ncu launch list of one eager step of the shipped tree.
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
(time timeout 900 python -m pytest tests/ -x -q -m gpu) > $O/gputest_final.log 2>&1; echo "rc=$?" >> $O/gputest_final.log; tail -5 $O/gputest_final.log
(time timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
(time timeout 400 python bench.py) > $O/bench_default_final.json 2> $O/bench_default_final.err; cut -c1-200 $O/bench_default_final.json
(time timeout 200 python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_reference_final.json 2> $O/bench_reference_final.err; cut -c1-200 $O/bench_reference_final.json
(time timeout 240 python scripts/overlap_check.py 32 23) > $O/overlap_variants.json 2> $O/overlap_variants.err; echo "rc=$?" >> $O/overlap_variants.err; cat $O/overlap_variants.json; tail -3 $O/overlap_variants.err
(time timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram.csv python scripts/profile_step.py 32) > $O/ncu_list.log 2>&1; tail -2 $O/ncu_list.log; wc -l $O/launches_dram.csv
