# Round profile (run under gpurun): ncu launch list of one eager step + full captures of the chained conv kernel (one forward, one
# input-gradient launch) and the batched weight gradient; summarise with scripts/summarize_launches.py / summarize_ncu.py into profiles/.
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/launches_r1g.csv python scripts/profile_step.py 32 > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_chain_kernel -s 68 -c 2 -f -o gpurun_out/r01g_conv_chain python scripts/profile_step.py 32 > gpurun_out/ncu_full_chain.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad9_tc_batched -s 10 -c 1 -f -o gpurun_out/r01g_wgrad9b python scripts/profile_step.py 32 > gpurun_out/ncu_full_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1g.csv
tail -3 gpurun_out/ncu_full_chain.log
