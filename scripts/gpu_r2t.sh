# GPU call T (8 GPUs): the driver's scaling sweep N = 1, 2, 4, 8 + DDP tests + 8-GPU tile inference
set -x
O=gpurun_out/r2t; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout -s KILL 600 python -m pytest tests/test_ddp_gpu.py -q -s > $O/ddp_tests.log 2>&1; echo "rc=$?" >> $O/ddp_tests.log; tail -n 3 $O/ddp_tests.log
timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4 8; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_n$n.json 2> $O/bench_n$n.err
done
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 8 --mode infer --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_infer_n8.json 2> $O/bench_infer_n8.err
cut -c1-230 $O/bench_n1.json $O/bench_n2.json $O/bench_n4.json $O/bench_n8.json $O/bench_infer_n8.json; tail -n 3 $O/*.err
