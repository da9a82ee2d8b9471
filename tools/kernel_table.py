#!/usr/bin/env python
"""Per-kernel evidence table from an ncu CSV holding, per launch, gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum and sm__pipe_tensor_cycles_active / sm__throughput percentages.
Usage: tools/kernel_table.py file.csv [hbm_peak_GBps]  -> markdown on stdout."""
import csv, re, sys
path = sys.argv[1]; peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6582.5
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
per = {}
for r in csv.DictReader(lines):
    k = r["ID"]
    d = per.setdefault(k, {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]; mname = r["Metric Name"]
    if mname.startswith("gpu__time_duration"):
        v = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0) * v
    if mname.startswith("dram__bytes"):
        v = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0) * v
    d[mname] = v
agg = {}
for d in per.values():
    short = re.sub(r"[<(].*", "", d["name"]).replace("void ", "").replace("yb::", "")
    if short.startswith("at::") or short.startswith("at_cuda") or "elementwise" in short:
        short = "(torch fill/copy: test harness buffers)"
    a = agg.setdefault(short, {"n": 0, "us": 0.0, "bytes": 0.0, "tc": 0.0, "sm": 0.0})
    t = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["us"] += t
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tc = [v for k, v in d.items() if k.startswith("sm__pipe_tensor") or k.startswith("sm__pipe_tc")]
    a["tc"] += (max(tc) if tc else 0.0) * t
    a["sm"] += d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * t
T = sum(a["us"] for a in agg.values())
print("| kernel | launches | total us | share | DRAM GB/s (r+w) | of HBM peak | tensor pipe % | SM busy % |")
print("|---|--:|--:|--:|--:|--:|--:|--:|")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    if a["us"] <= 0: continue
    gbs = a["bytes"] / a["us"] / 1e3
    print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100*a['us']/T:.1f}% | {gbs:.0f} | {gbs/peak:.2f} | {a['tc']/a['us']:.1f} | {a['sm']/a['us']:.1f} |")
print(f"\ntotal {T:.1f} us over {sum(a['n'] for a in agg.values())} launches (ncu-serialised, cold caches per launch); HBM peak {peak} GB/s (MEASURED_PEAKS.json)")
