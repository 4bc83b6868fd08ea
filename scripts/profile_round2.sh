# Round-2 profile (run under gpurun): (1) launch list of one eager step with duration + DRAM bytes per launch, (2) ncu --set full
# captures of the resident dense-block kernel (one forward, one input-gradient launch) and the batched weight gradient.
# Summaries: scripts/summarize_launches.py, scripts/summarize_hbm.py, scripts/summarize_ncu.py -> profiles/r02_*.md
set -x
O=gpurun_out/r2prof; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram.csv python scripts/profile_step.py 32 > $O/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rdb_resident_kernel -s 34 -c 1 -f -o $O/r02_rdb_fwd python scripts/profile_step.py 32 > $O/ncu_full_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rdb_resident_kernel -s 103 -c 1 -f -o $O/r02_rdb_dgrad python scripts/profile_step.py 32 > $O/ncu_full_dgrad.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad9_tc_batched -s 10 -c 1 -f -o $O/r02_wgrad9b python scripts/profile_step.py 32 > $O/ncu_full_wgrad.log 2>&1
ls -la $O; tail -3 $O/ncu_full_fwd.log
