# GPU call X: how much does programmatic dependent launch buy on the conv launches (to size extending it to the helper kernels)
set -x
O=gpurun_out/r2x; mkdir -p $O
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json
SSR_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_nopdl.json 2> $O/bench_nopdl.err; cut -c1-200 $O/bench_nopdl.json
