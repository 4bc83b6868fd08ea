# Round-2 (second half) profile, run under gpurun: (1) launch list of one eager step with duration + DRAM bytes per launch,
# (2) ncu --set full captures: the resident dense-block kernel with an RRDB (three blocks) per launch -- one forward, one
# input-gradient launch --, the largest bilinear x2 launch and the generator's weight-pack launch.
# Summaries: scripts/summarize_launches.py, scripts/summarize_hbm.py, scripts/summarize_ncu.py -> profiles/r02b_*.md
set -x
O=gpurun_out/r2bprof; mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram.csv python scripts/profile_step.py 32 > $O/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rdb_resident_kernel -s 11 -c 1 -f -o $O/r02b_rdb_fwd python scripts/profile_step.py 32 > $O/ncu_full_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rdb_resident_kernel -s 34 -c 1 -f -o $O/r02b_rdb_dgrad python scripts/profile_step.py 32 > $O/ncu_full_dgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:upsample_bilinear2x_kernel -s 2 -c 1 -f -o $O/r02b_bilinear python scripts/profile_step.py 32 > $O/ncu_full_bilinear.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pack_batched_kernel -c 1 -f -o $O/r02b_pack python scripts/profile_step.py 32 > $O/ncu_full_pack.log 2>&1
ls -la $O; tail -3 $O/ncu_full_fwd.log
