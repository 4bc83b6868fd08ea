"""`ncu --set full` report -> markdown summary (key metrics per captured launch + top stall sites).  Usage:
python scripts/summarize_ncu.py gpurun_out/r01_conv_tc.ncu-rep profiles/r01_ncu_conv_tc.md"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = csv.reader(raw.splitlines())
hdr, units = next(rd), next(rd)
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem/CTA")]
idx = [(lbl, hdr.index(k)) for k, lbl in want if k in hdr]
lines = [f"# {rep}: ncu --set full --clock-control none (cold cache, serialised: use shares / ratios)\n",
         "| " + " | ".join(l for l, _ in idx) + " |", "|" + "---|" * len(idx)]
for row in rd:
    cells = []
    for lbl, i in idx:
        v = row[i]
        if lbl == "kernel":
            v = v.split("(")[0].replace("void ", "")
        elif units[i] and lbl not in ("grid",):
            try:
                v = f"{float(v):.3g} {units[i]}"
            except ValueError:
                pass
        cells.append(v)
    lines.append("| " + " | ".join(cells) + " |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, rows, h = 0, {}, None
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        kern += 1
        h = None
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if h is None:
        continue
    rows.setdefault(kern, (h, []))[1].append(r)
for k, (h, rr) in rows.items():
    i_s, i_src = h.index("# Samples"), h.index("Source")
    tot = sum(int(r[i_s]) for r in rr) or 1
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = sorted(((h[i], sum(int(r[i] or 0) for r in rr)) for i in stall_cols), key=lambda kv: -kv[1])[:5]
    lines.append(f"\n## launch {k}: warp-state samples {tot}; " + ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in agg))
    lines.append("\n| samples | SASS |\n|---:|---|")
    for r in sorted(rr, key=lambda r: -int(r[i_s]))[:8]:
        lines.append(f"| {r[i_s]} | `{r[i_src].strip()[:80]}` |")
    if k >= 3:
        break
open(out, "w").write("\n".join(lines) + "\n")
# mean DRAM traffic per captured launch -> profiles/ncu_traffic.json (bench.py's roofline.traffic)
import json, os
try:
    raw2 = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd2 = csv.reader(raw2.splitlines())
    h2, u2 = next(rd2), next(rd2)
    ir, iw, ik = h2.index("dram__bytes_read.sum"), h2.index("dram__bytes_write.sum"), h2.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, kname = 0.0, 0, None
    for row in rd2:
        tot += float(row[ir]) * scale.get(u2[ir], 1) + float(row[iw]) * scale.get(u2[iw], 1)
        n += 1
        kname = row[ik].split("(")[0].replace("void ", "").split("<")[0].replace("ssr::", "")
    jpath = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    data = json.load(open(jpath)) if os.path.exists(jpath) else {}
    data[kname] = {"dram_bytes_per_launch": tot / max(n, 1), "launches_captured": n, "report": os.path.basename(rep),
                   "note": "mean dram__bytes_read.sum + dram__bytes_write.sum over the captured launches (cold cache, ncu --set full)"}
    json.dump(data, open(jpath, "w"), indent=1)
except Exception as e:  # noqa
    print("traffic json not updated:", e)
print("\n".join(lines[:12]))
