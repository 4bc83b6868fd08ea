# GPU call K: ncu --set full of the plain 64 -> 64 @ 128^2 conv (what bounds it?), tests + bench with the epilogue bias gradients
set -x
O=gpurun_out/r2k; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o $O/r02_conv64 python scripts/bench_conv.py 32 "hr 128" > $O/ncu_conv64.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_depth_parity_gpu.py --deselect tests/test_conv_tc_gpu.py > $O/tests.log 2>&1
tail -n 3 $O/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err
cut -c1-200 $O/bench.json
ls -la $O
