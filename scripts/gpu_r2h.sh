# 2-GPU call: DDP equivalence + rank-dependent-init tests, bench at N = 1 and N = 2 on the same box (all-reduce overlap)
set -x
O=gpurun_out/r2h; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
timeout -s KILL 900 python -m pytest tests/test_ddp_gpu.py -q -s > $O/ddp_tests.log 2>&1; echo "rc=$?" >> $O/ddp_tests.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --mode infer --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_infer_n2.json 2> $O/bench_infer_n2.err
timeout -s KILL 300 python bench.py --gpus 1 --mode infer --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_infer_n1.json 2> $O/bench_infer_n1.err
tail -5 $O/ddp_tests.log; cut -c1-250 $O/bench_n1.json $O/bench_n2.json $O/bench_infer_n1.json $O/bench_infer_n2.json; tail -3 $O/*.err
