#!/bin/bash
# round-2 GPU session 1: tests per file (own process each), epilogue A/B probes, bench, reference arm x2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
for f in tests/test_gpu_conv.py tests/test_gpu_train_ops.py tests/test_gpu_path.py; do
  timeout 900 python -m pytest $f -m gpu -q --maxfail=12 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_1_$(basename $f .py).log
  tail -3 gpurun_out/r02_1_$(basename $f .py).log
done
echo "=== probes"
P=gpurun_out/r02_1_probes.txt; : > $P
probe() { timeout 120 python tools/conv_probe.py "$@" >> $P 2>&1; }
for epi in reg tma; do
  export YB_CONV_EPI=$epi; [ $epi = tma ] && unset YB_CONV_EPI
  for dbg in 0 7; do
    export YB_CONV_DBG=$dbg
    probe 64 52 52 256 128 1 1
    probe 64 26 26 512 256 1 1
    probe 64 13 13 1024 512 1 1
    probe 64 104 104 128 64 1 1
  done
  export YB_CONV_DBG=0
  probe 64 208 208 64 32 1 1
  probe 64 52 52 128 256 3 1 10 res
  probe 64 26 26 256 512 3 1 10 res
  probe 64 13 13 512 1024 3 1 10 res
  probe 64 104 104 64 128 3 1 10 res
  probe 64 208 208 32 64 3 1 10 res
done
unset YB_CONV_EPI; export YB_CONV_DBG=0
export YB_CONV_BRES=1
probe 64 52 52 256 128 1 1
probe 64 104 104 128 64 1 1
probe 64 208 208 64 32 1 1
unset YB_CONV_BRES
export YB_CONV_MODE=1cta
probe 64 26 26 512 256 1 1
probe 64 13 13 1024 512 1 1
export YB_CONV_MODE=2cta
probe 64 52 52 256 128 1 1
unset YB_CONV_MODE YB_CONV_DBG
cat $P
echo "=== bench"
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02_1_bench.json 2> gpurun_out/r02_1_bench.err; tail -c 600 gpurun_out/r02_1_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_1_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "train", d["train"]["ms_per_step"], "train608", d.get("train608",{}).get("ms_per_step"), "lat", d["latency_batch1"]["ms_median"])
except Exception as e: print("bench parse failed", e)
PY
echo "=== reference arm x2"
for i in 1 2; do timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_1_ref$i.json 2> gpurun_out/r02_1_ref$i.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_1_ref$i.json').read().strip().splitlines()[-1]);print(d['value'],[ (c['workers'],c['threads'],c.get('images_per_s'),c.get('skipped')) for c in d['cpu_baseline']['candidates']])"; done
