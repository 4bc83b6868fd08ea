"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid) -> profiles/*.md"""
import csv
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
per_kernel = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    grid = r.get("Grid Size", "")
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        val *= 1e3
    elif unit in ("ms", "msecond"):
        val *= 1e6
    agg[(name, grid)][0] += 1
    agg[(name, grid)][1] += val
    per_kernel[name][0] += 1
    per_kernel[name][1] += val
    total += val
with open(out, "w") as fh:
    fh.write(f"# launch list summary of one training step ({path})\n\n")
    fh.write("ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    fh.write(f"total {total / 1e6:.2f} ms over {sum(v[0] for v in per_kernel.values())} launches\n\n")
    fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for name, (n, t) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| {name} | {n} | {t / 1e6:.3f} | {100 * t / total:.1f}% |\n")
    fh.write("\n## by (kernel, grid)\n\n| kernel | grid | launches | total ms | avg us | share |\n|---|---|---:|---:|---:|---:|\n")
    for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        fh.write(f"| {name} | {grid} | {n} | {t / 1e6:.3f} | {t / n / 1e3:.1f} | {100 * t / total:.1f}% |\n")
print(open(out).read()[:6000])
