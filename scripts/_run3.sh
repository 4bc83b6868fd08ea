set -x
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_ddp_gpu.py 2>&1 | tail -6
timeout 200 python scripts/time_g.py 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
SSR_CHAIN_TIMELINE=1 timeout 100 python scripts/chain_timeline.py > gpurun_out/chain_timeline.txt 2>&1; grep -E "^==|mean over|->|layer total" gpurun_out/chain_timeline.txt
