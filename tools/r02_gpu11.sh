#!/bin/bash
# round-2 GPU session 11: consolidated configuration (TMA epilogue, EG=2 on 1x1, halo on Cin=32): all GPU tests + bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -15 > gpurun_out/r02_11_tests.log; tail -8 gpurun_out/r02_11_tests.log | cut -c1-300
echo "=== bench"
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02_11_bench.json 2> gpurun_out/r02_11_bench.err; tail -c 600 gpurun_out/r02_11_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_11_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "conv ms", r["ms_per_step_conv"], "stem", r["ms_per_step_stem"], "nms", r["ms_per_step_nms"], "unfused", d["unfused_api_ms_per_step"], "train", d["train"]["ms_per_step"], "train608", d.get("train608",{}).get("ms_per_step"), "lat", d["latency_batch1"]["ms_median"])
except Exception as e: print("bench parse failed", e)
PY
