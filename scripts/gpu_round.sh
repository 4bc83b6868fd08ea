#!/bin/bash
# Final visit of the round: the whole GPU suite exactly as the driver runs it, smoke(), and the default bench line.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/ -x -q -m gpu) > $O/gputest_final.log 2>&1; echo "rc=$?" >> $O/gputest_final.log; tail -5 $O/gputest_final.log
(time timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
(time timeout 400 python bench.py) > $O/bench_default_final.json 2> $O/bench_default_final.err; cut -c1-200 $O/bench_default_final.json
(time timeout 200 python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_reference_final.json 2> $O/bench_reference_final.err; cut -c1-300 $O/bench_reference_final.json
