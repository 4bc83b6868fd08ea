# GPU call R: full GPU suite incl. depth parity, smoke, default bench
set -x
O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_all.log 2>&1; tail -n 4 $O/tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-400 $O/bench_default.json
