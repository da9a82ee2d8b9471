#!/bin/bash
# round-2 GPU session 16: stem fused into Conv_1 in the plan: timing, all GPU tests, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash tools/r02_gpu14.sh 2>&1 | tail -5
echo "=== tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -15 > gpurun_out/r02_16_tests.log; tail -6 gpurun_out/r02_16_tests.log | cut -c1-300
echo "=== bench"
for fuse in 1 0; do
YB_STEM_FUSE=$fuse timeout 900 python bench.py --no-cpu-baseline --steps 20 --no-train > gpurun_out/r02_16_bench_fuse$fuse.json 2> gpurun_out/r02_16_bench.err; tail -c 400 gpurun_out/r02_16_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_16_bench_fuse$fuse.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("fuse=$fuse value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "conv ms", r["ms_per_step_conv"], "stem", r["ms_per_step_stem"], "nms", r["ms_per_step_nms"], "lat", d["latency_batch1"]["ms_median"], d["latency_batch1"]["cuda_graph_ms_median"])
except Exception as e: print("bench parse failed", e)
PY
done
