set -x
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_ddp_gpu.py 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
