#!/bin/bash
# BN statistics on the TMA epilogue: tests, step A/B; synccheck on a standalone launch with one epilogue group / the pair kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -8 > gpurun_out/r02_i_tests.log; tail -4 gpurun_out/r02_i_tests.log | cut -c1-300
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_CONV_STAT_TMA=0" "" > gpurun_out/r02_i_train_ab.txt 2>&1; cat gpurun_out/r02_i_train_ab.txt | cut -c1-200
echo "=== synccheck: standalone 1x1 + residual launch"
for cfg in "YB_CONV_EG=1" "YB_CONV_MODE=2cta" "YB_CONV_EPI=reg"; do
  echo "--- [$cfg]"
  env $cfg timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python tools/conv_probe.py 2 26 26 256 256 1 1 1 res 2>&1 | grep -E "median|ERROR SUMMARY|Barrier error|by thread|located|Device Frame.*kernel" | head -8
done > gpurun_out/r02_i_synccheck.txt 2>&1; cat gpurun_out/r02_i_synccheck.txt | cut -c1-220
