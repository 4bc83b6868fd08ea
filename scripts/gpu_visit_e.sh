#!/bin/bash
# last seconds of the round's GPU budget: the clocks sampler against the real nvidia-smi, then the short bench line
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 8 python - > $O/sampler_check.log 2>&1 <<'PY'
import importlib.util, time
spec = importlib.util.spec_from_file_location("bench", "bench.py"); b = importlib.util.module_from_spec(spec)
import sys; sys.argv = ["bench.py"]; spec.loader.exec_module(b)
s = b.ClockSampler(0); t0 = time.time(); s.start(); time.sleep(1.5); n_mid = len(s.rows); s.stop_flag = True; s.join(3)
print("rows after 1.5 s:", n_mid, "summary:", s.summary(), "alive:", s.is_alive())
PY
cat $O/sampler_check.log
(timeout 40 python bench.py --no-cpu-baseline --no-extras) > $O/bench_shipped2_n1.json 2> $O/bench_shipped2_n1.err; cut -c1-200 $O/bench_shipped2_n1.json; python -c "
import json; d=json.loads(open('$O/bench_shipped2_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'], d['clocks'])"
