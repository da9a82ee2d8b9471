#!/bin/bash
# round-2 GPU session 5: staging-tile + coalesced-store epilogue: conv tests, traces, probes, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_train_ops.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/r02_5_tests.log; tail -3 gpurun_out/r02_5_tests.log
timeout 900 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider --tb=short -k "detect or forward_matches or nms_bit_exact" 2>&1 | tail -30 > gpurun_out/r02_5_tests2.log; tail -3 gpurun_out/r02_5_tests2.log
echo "=== traces"
T=gpurun_out/r02_5_traces.txt; : > $T
trace() { timeout 120 python tools/conv_trace.py "$@" 2>&1 | head -9 >> $T; }
trace 64 52 52 256 128 1 1
YB_CONV_EPI=tma trace 64 52 52 256 128 1 1
YB_CONV_EPI=tma YB_CONV_DBG=7 trace 64 52 52 256 128 1 1
YB_CONV_DBG=7 trace 64 52 52 256 128 1 1
YB_CONV_BRES=1 trace 64 52 52 256 128 1 1
trace 64 104 104 128 64 1 1
YB_CONV_MODE=1cta trace 64 52 52 128 256 3 1 res
trace 64 208 208 32 64 3 1 res
cut -c1-250 $T
echo "=== probes"
P=gpurun_out/r02_5_probes.txt; : > $P
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
probe 64 208 208 32 64 3 2
probe 64 104 104 64 128 3 2
YB_CONV_MODE=1cta probe 64 26 26 512 256 1 1
YB_CONV_MODE=1cta probe 64 13 13 1024 512 1 1
YB_CONV_MODE=2cta probe 64 52 52 256 128 1 1
YB_CONV_EPI=tma probe 64 52 52 256 128 1 1
YB_CONV_EPI=tma probe 64 26 26 512 256 1 1
YB_CONV_EPI=tma probe 64 52 52 128 256 3 1 10 res
YB_CONV_EPI=tma probe 64 26 26 256 512 3 1 10 res
YB_CONV_EPI=reg probe 64 52 52 256 128 1 1
YB_CONV_EPI=reg probe 64 208 208 64 32 1 1
YB_CONV_BRES=1 probe 64 52 52 256 128 1 1
YB_CONV_BRES=1 probe 64 104 104 128 64 1 1
YB_CONV_BRES=1 probe 64 208 208 64 32 1 1
YB_CONV_BRES=1 probe 64 208 208 32 64 3 1 10 res
YB_CONV_BRES=1 probe 64 104 104 64 128 3 1 10 res
cat $P
echo "=== bench"
timeout 900 python bench.py --no-cpu-baseline --steps 10 --no-train608 > gpurun_out/r02_5_bench.json 2> gpurun_out/r02_5_bench.err; tail -c 800 gpurun_out/r02_5_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_5_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "conv ms", r["ms_per_step_conv"], "stem", r["ms_per_step_stem"], "nms", r["ms_per_step_nms"], "unfused", d["unfused_api_ms_per_step"], "train", d["train"]["ms_per_step"], "lat", d["latency_batch1"]["ms_median"], "det", d["detections_per_step"])
except Exception as e: print("bench parse failed", e)
PY
