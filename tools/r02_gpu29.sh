#!/bin/bash
# detection heads 1-2 on a side stream: tests, A/B at batch 64 and batch 1, then the bench (with the faster setting)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 > gpurun_out/r02_t_tests.log; tail -4 gpurun_out/r02_t_tests.log | cut -c1-300
echo "=== infer A/B"
timeout 600 python tools/infer_ab.py 64 416 12 -- "" "YB_HEAD_STREAM=0" "" "YB_HEAD_STREAM=0" > gpurun_out/r02_t_infer_ab.txt 2>&1
timeout 300 python tools/infer_ab.py 1 416 40 -- "" "YB_HEAD_STREAM=0" >> gpurun_out/r02_t_infer_ab.txt 2>&1
cat gpurun_out/r02_t_infer_ab.txt | cut -c1-200
SET=$(python - <<'PY'
import re
on=[];off=[]
for l in open("gpurun_out/r02_t_infer_ab.txt"):
    m=re.match(r"\[(.*?)\] batch 64 .* median ([\d.]+) ms", l)
    if m: (off if "HEAD_STREAM=0" in m.group(1) else on).append(float(m.group(2)))
print("0" if (on and off and min(on) > min(off) * 0.997) else "")
PY
)
echo "bench with YB_HEAD_STREAM='$SET'"
echo "=== bench"
YB_HEAD_STREAM=$SET timeout 1500 python bench.py > gpurun_out/r02_t_bench.json 2> gpurun_out/r02_t_bench.err; tail -c 300 gpurun_out/r02_t_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_t_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]; c=d.get("cpu_baseline",{})
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "train", d["train"]["ms_per_step"], d["train"]["e2e"]["ms_per_step"], "train608", d.get("train608",{}).get("ms_per_step"), "lat", d["latency_batch1"]["ms_median"], d["latency_batch1"]["cuda_graph_ms_median"], "cpu", c.get("value"), "clocks", d["clocks"])
except Exception as e: print("bench parse failed", e)
PY
