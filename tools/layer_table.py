#!/usr/bin/env python
"""Per-layer table of ONE inference step from an ncu CSV of tools/infer_probe.py (metrics: gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum, sm__pipe_tensor_cycles_active..., sm__throughput...): the launches are the
stem, the 74 tensor-core convs in creation order, nms_select, nms_gather.  Usage: tools/layer_table.py file.csv [batch] [size]
-> markdown on stdout and profiles/conv_traffic.json-style totals on stderr."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200.model import yolov3
path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64; S = int(sys.argv[3]) if len(sys.argv) > 3 else 416
PEAK_TF, PEAK_BW = 1371.6, 6582.5
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
per, order = {}, []
for r in csv.DictReader(lines):
    k = r["ID"]
    if k not in per:
        per[k] = {"name": r["Kernel Name"]}; order.append(k)
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u, mname = r["Metric Unit"], r["Metric Name"]
    if mname.startswith("gpu__time_duration"):
        v = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0) * v
    if mname.startswith("dram__bytes"):
        v = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0) * v
    per[k][mname] = v
launches = [per[k] for k in order if "at::" not in per[k]["name"] and "elementwise" not in per[k]["name"]]
table = yolov3.conv_table(80)
# with the stem fused into Conv_1 (YB_STEM_FUSE, the default) launch 0 is conv_halo_kernel<.., STEMW > 0, ..> and covers layers 0 + 1
m0 = re.search(r"conv_halo_kernel<[^,]+,\s*\d+,\s*\d+,\s*\d+,\s*(\d+)", launches[0]["name"])
fused = bool(m0 and int(m0.group(1)) > 0)
nl = 76 if fused else 77
launches = launches[:nl]
print(f"# per-layer table of one inference step (batch {B} @{S}, fp16), from {os.path.basename(path)}\n")
print("ncu-serialised durations with cold caches (bench.py times the whole step warm).  TFLOP/s = 2*pixels*k*k*cin*cout / time; peaks: "
      f"{PEAK_TF} TFLOP/s (sustained bf16, measured), {PEAK_BW} GB/s (measured).\n")
print("| layer | conv | out HxW | kernel | us | TFLOP/s | of peak | DRAM MB | DRAM GB/s | of HBM peak | tensor pipe % |")
print("|--:|---|---|---|--:|--:|--:|--:|--:|--:|--:|")
div, tot_us, conv_us, conv_bytes, conv_flop = 1, 0.0, 0.0, 0.0, 0.0
# spatial size per layer: walk the table
h = S
sizes = []
c_prev = None
hs = S
for i, (cin, cout, k, s, bn) in enumerate(table):
    pass
def layer_hw():
    out = []
    hh = S
    idx = 0
    def conv(s=1):
        nonlocal hh
        hh //= s
        out.append(hh)
    conv(); conv(2)
    for reps, f in ((1, 32), (2, 64), (8, 128), (8, 256), (4, 512)):
        for _ in range(reps):
            conv(); conv()
        if f != 512:
            conv(2)
    for g in (S // 32, S // 16, S // 8):
        hh = g
        for _ in range(7):
            conv()
        if g != S // 8:
            out.append(g)          # the upsampling 1x1 runs at the coarse resolution
    return out
hw = layer_hw()
for li, L in enumerate(launches):
    i = li + 1 if (fused and li > 0) else li
    name = re.sub(r"\(.*", "", L["name"]).replace("void ", "").replace("yb::", "")
    us = L.get("gpu__time_duration.sum", 0.0); by = L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
    tp = L.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    tot_us += us
    if fused and li == 0:
        fl = sum(2.0 * B * hw[j] * hw[j] * table[j][2] ** 2 * table[j][0] * table[j][1] for j in (0, 1))
        tf = fl / us / 1e6
        conv_us += us; conv_bytes += by; conv_flop += fl
        print(f"| 0+1 | 3x3/1 3->32 + 3x3/2 32->64 (fused) | {hw[1]}x{hw[1]} | `{name[:46]}` | {us:.1f} | {tf:.0f} | {tf / PEAK_TF:.2f} | {by / 1e6:.0f} | {by / us / 1e3:.0f} | {by / us / 1e3 / PEAK_BW:.2f} | {tp:.0f} |")
    elif i < 75:
        cin, cout, k, s, bn = table[i]
        o = hw[i]
        fl = 2.0 * B * o * o * k * k * cin * cout
        tf = fl / us / 1e6
        if i >= 1:
            conv_us += us; conv_bytes += by; conv_flop += fl
        print(f"| {i} | {k}x{k}/{s} {cin}->{cout} | {o}x{o} | `{name[:46]}` | {us:.1f} | {tf:.0f} | {tf / PEAK_TF:.2f} | {by / 1e6:.0f} | {by / us / 1e3:.0f} | {by / us / 1e3 / PEAK_BW:.2f} | {tp:.0f} |")
    else:
        print(f"| - | - | - | `{name[:46]}` | {us:.1f} | - | - | {by / 1e6:.0f} | {by / us / 1e3:.0f} | {by / us / 1e3 / PEAK_BW:.2f} | - |")
print(f"\nsum {tot_us:.0f} us; convs ({'layers 0..74, stem fused into Conv_1' if fused else 'layers 1..74'}): {conv_us:.0f} us, {conv_flop / conv_us / 1e6:.0f} TFLOP/s, DRAM {conv_bytes / 1e9:.2f} GB")
sys.stderr.write(json.dumps({"batch": B, "size": S, "dram_bytes_per_step": conv_bytes, "conv_us_ncu": conv_us, "source": os.path.basename(path)}) + "\n")
