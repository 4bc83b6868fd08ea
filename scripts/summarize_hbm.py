"""`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of one training step ->
per-kernel achieved HBM GB/s against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs), profiles/*.md.
Per-launch numbers under ncu are cold-cache and serialised: DRAM bytes are exact per launch, durations are upper bounds.

    python scripts/summarize_hbm.py gpurun_out/r2/launches_dram.csv profiles/r02_hbm_table.md
"""
import csv
import json
import os
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6581.6
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
scale_t = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}
scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per_id = defaultdict(dict)
for r in csv.DictReader(lines):
    m = r.get("Metric Name")
    if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    v *= scale_t.get(u, 1.0) if m.startswith("gpu__time") else scale_b.get(u, 1.0)
    per_id[r["ID"]]["name"] = r["Kernel Name"].split("(")[0].replace("void ", "")
    per_id[r["ID"]][m] = v
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per_id.values():
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
total_t = sum(a[1] for a in agg.values())
TENSOR = ("conv_tc_kernel", "conv_chain_kernel", "rdb_resident_kernel", "wgrad")
with open(out, "w") as fh:
    fh.write(f"# DRAM traffic and achieved HBM bandwidth per kernel, one training step ({os.path.basename(path)})\n\n")
    fh.write(f"peak = {peak:.1f} GB/s (measured copy bandwidth, MEASURED_PEAKS.json).  ncu serialises launches and starts each from a cold "
             "L2, so `GB/s` = DRAM bytes / duration is what the kernel achieves on its own; tensor-core kernels are listed for their traffic only.\n\n")
    fh.write("| kernel | launches | ms | share | DRAM read MB | DRAM write MB | GB/s | of peak | bound |\n|---|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for name, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (rd + wr) / t if t > 0 else 0.0     # bytes / ns = GB/s
        bound = "tensor" if any(k in name for k in TENSOR) else "hbm"
        fh.write(f"| {name} | {n} | {t / 1e6:.3f} | {100 * t / total_t:.1f}% | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f}% | {bound} |\n")
print(open(out).read()[:5000])
