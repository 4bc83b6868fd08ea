# GPU call S: bilinear kernels (one thread per input pixel)
set -x
O=gpurun_out/r2s; mkdir -p $O
timeout 600 python -m pytest tests/test_resample_gpu.py tests/test_modules_gpu.py tests/test_train_gpu.py -x -q > $O/tests.log 2>&1; tail -n 4 $O/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench2.json 2> $O/bench2.err; cut -c1-200 $O/bench2.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:bilinear -c 18 --csv --log-file $O/bilinear.csv python scripts/profile_step.py 32 > $O/ncu_list.log 2>&1
grep -c bilinear $O/bilinear.csv
