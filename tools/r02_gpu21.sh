#!/bin/bash
# synccheck "Missing init" classification: (a) residual barriers first in the barrier block, (b) which blocks are flagged on a 338-tile launch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "YB_CONV_DBG=0" "YB_CONV_DBG=128"; do
  echo "--- [$cfg] 2 26 26 256 128 (11 tiles)"
  env $cfg timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python tools/conv_probe.py 2 26 26 256 128 1 1 1 res 2>&1 | grep -E "median|ERROR SUMMARY|Barrier error|by thread|located" | head -12
done > gpurun_out/r02_j_synccheck.txt 2>&1
echo "--- 64 26 26 256 128 (338 tiles on 148 CTAs): flagged blocks" >> gpurun_out/r02_j_synccheck.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 100000 python tools/conv_probe.py 64 26 26 256 128 1 1 0 res 2>&1 | grep -E "by thread|ERROR SUMMARY|located" | sed -E 's/by thread \([0-9]+,0,0\) in //' | sort | uniq -c | sort -rn | head -12 >> gpurun_out/r02_j_synccheck.txt 2>&1
cat gpurun_out/r02_j_synccheck.txt | cut -c1-200
