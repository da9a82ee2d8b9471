#!/bin/bash
# round-2 GPU session 3: all GPU tests, traces after the rolled-epilogue rewrite, probes, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
for f in tests/test_gpu_conv.py tests/test_gpu_train_ops.py tests/test_gpu_path.py; do
  timeout 900 python -m pytest $f -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -80 > gpurun_out/r02_3_$(basename $f .py).log
  tail -4 gpurun_out/r02_3_$(basename $f .py).log
done
echo "=== traces"
T=gpurun_out/r02_3_traces.txt; : > $T
trace() { timeout 120 python tools/conv_trace.py "$@" 2>&1 | head -14 >> $T; }
trace 64 52 52 256 128 1 1
YB_CONV_MODE=1cta trace 64 26 26 512 256 1 1
trace 64 104 104 128 64 1 1
YB_CONV_MODE=1cta trace 64 52 52 128 256 3 1 res
trace 64 208 208 32 64 3 1 res
cut -c1-250 $T
echo "=== probes"
P=gpurun_out/r02_3_probes.txt; : > $P
probe() { timeout 120 python tools/conv_probe.py "$@" >> $P 2>&1; }
probe 64 52 52 256 128 1 1
probe 64 26 26 512 256 1 1
probe 64 13 13 1024 512 1 1
probe 64 104 104 128 64 1 1
probe 64 208 208 64 32 1 1
probe 64 52 52 128 256 3 1 10 res
probe 64 26 26 256 512 3 1 10 res
probe 64 13 13 512 1024 3 1 10 res
probe 64 104 104 64 128 3 1 10 res
probe 64 208 208 32 64 3 1 10 res
probe 64 52 52 128 256 3 1
YB_CONV_MODE=1cta probe 64 26 26 512 256 1 1
YB_CONV_MODE=1cta probe 64 13 13 1024 512 1 1
YB_CONV_MODE=2cta probe 64 52 52 256 128 1 1
YB_CONV_EPI=reg probe 64 52 52 256 128 1 1
YB_CONV_EPI=reg probe 64 208 208 64 32 1 1
cat $P
echo "=== bench"
timeout 900 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_3_bench.json 2> gpurun_out/r02_3_bench.err; tail -c 800 gpurun_out/r02_3_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_3_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "conv ms", r["ms_per_step_conv"], "stem", r["ms_per_step_stem"], "nms", r["ms_per_step_nms"], "unfused", d["unfused_api_ms_per_step"], "train", d["train"]["ms_per_step"], "train608", d.get("train608",{}).get("ms_per_step"), "lat", d["latency_batch1"]["ms_median"], "det", d["detections_per_step"])
except Exception as e: print("bench parse failed", e)
PY
