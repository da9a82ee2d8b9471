#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "YB_CONV_DBG=256" "YB_CONV_DBG=512" "YB_CONV_DBG=768"; do
  echo "--- [$cfg] 2 26 26 256 128 (11 tiles)"
  env $cfg timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python tools/conv_probe.py 2 26 26 256 128 1 1 1 res 2>&1 | grep -E "median|ERROR SUMMARY|Barrier error|by thread|located" | head -8
done > gpurun_out/r02_k_synccheck.txt 2>&1
cat gpurun_out/r02_k_synccheck.txt | cut -c1-200
