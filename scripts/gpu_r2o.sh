# GPU call O: launch list of one eager step (durations + DRAM bytes) with the current kernels
set -x
O=gpurun_out/r2o; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $O/launches_dram.csv python scripts/profile_step.py 32 > $O/ncu_list.log 2>&1
tail -n 2 $O/ncu_list.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o $O/r02_conv64_halo python scripts/bench_conv_big.py "VGG conv1_2" > $O/ncu_conv64_halo.log 2>&1
ls -la $O
