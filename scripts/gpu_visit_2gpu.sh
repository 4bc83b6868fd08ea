#!/bin/bash
# 2-GPU check of the side lane: the DDP equivalence tests and the N = 2 bench line (one graph per phase around the asynchronous all-reduces).
# NOT run in round 2c (the GPU budget went to the 1-GPU visits; the per-phase capture is covered on one GPU by tests/test_overlap_gpu.py).
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
(time timeout 150 python -m pytest tests/test_ddp_gpu.py -x -q -s) > $O/ddp_tests_2gpu.log 2>&1; echo "rc=$?" >> $O/ddp_tests_2gpu.log; tail -6 $O/ddp_tests_2gpu.log
(time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras) > $O/bench_n2.json 2> $O/bench_n2.err; cut -c1-250 $O/bench_n2.json; tail -3 $O/bench_n2.err
(time timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras) > $O/bench_n1_2gpu_box.json 2> $O/bench_n1_2gpu_box.err; cut -c1-250 $O/bench_n1_2gpu_box.json
