# round-2 GPU call A: baseline + SBO probe + resident chain + TMEM dgrad at B=32 + depth parity
set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_depth_parity_gpu.py > $O/tests_default.log 2>&1; echo "rc=$?" >> $O/tests_default.log
SSR_CONV_TW=8 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -x -q > $O/probe_tw8.log 2>&1; echo "rc=$?" >> $O/probe_tw8.log
SSR_CONV_TW=8 SSR_DBG_PITCH=2 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -q > $O/probe_pitch2.log 2>&1; echo "rc=$?" >> $O/probe_pitch2.log
SSR_CONV_RESIDENT=1 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -q -k "chain_equals_plain" > $O/resident_fwd.log 2>&1; echo "rc=$?" >> $O/resident_fwd.log
SSR_CONV_RESIDENT=1 SSR_CHAIN_TIMELINE=1 timeout -s KILL 120 python scripts/chain_timeline.py > $O/resident_timeline.log 2>&1; echo "rc=$?" >> $O/resident_timeline.log
SSR_CONV_RESIDENT=1 timeout -s KILL 300 python -m pytest tests/test_conv_tc_gpu.py -q -k "chain_acc" > $O/resident_acc.log 2>&1; echo "rc=$?" >> $O/resident_acc.log
SSR_DGRAD_TMEM=1 timeout -s KILL 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py tests/test_modules_gpu.py -q > $O/dgrad_tmem.log 2>&1; echo "rc=$?" >> $O/dgrad_tmem.log
timeout -s KILL 600 python -m pytest tests/test_depth_parity_gpu.py -q -s > $O/depth_parity.log 2>&1; echo "rc=$?" >> $O/depth_parity.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
SSR_DGRAD_TMEM=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_tmem.json 2> $O/bench_tmem.err
SSR_DGRAD_TMEM=1 SSR_CONV_RESIDENT=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_tmem_res.json 2> $O/bench_tmem_res.err
tail -3 $O/*.log; cat $O/bench_*.json | cut -c1-300
