set -x
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -k "chain or planar" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_generator_gpu.py tests/test_train_gpu.py -x -q 2>&1 | tail -8
timeout 200 python scripts/time_g.py 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | cut -c1-300
