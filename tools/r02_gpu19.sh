#!/bin/bash
# wgrad split selection + loss fast path: training tests, step A/B over the epilogue-cost constant; synccheck on a standalone launch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests (training ops + path)"
timeout 1200 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_path.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -8 > gpurun_out/r02_h_tests.log; tail -4 gpurun_out/r02_h_tests.log | cut -c1-300
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_WGRAD_EPI=20" "YB_WGRAD_EPI=80" "YB_WGRAD_EPI=160" "YB_WGRAD_EPI=0" "" > gpurun_out/r02_h_train_ab.txt 2>&1; cat gpurun_out/r02_h_train_ab.txt | cut -c1-200
echo "=== synccheck: one standalone 1x1 + residual launch, 11 and 22 tiles"
for shp in "2 26 26 64 128" "4 26 26 64 128" "2 26 26 256 128"; do
  echo "--- [$shp]"
  timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python tools/conv_probe.py $shp 1 1 1 res 2>&1 | grep -E "median|ERROR SUMMARY|Barrier error|by thread|located" | head -8
done > gpurun_out/r02_h_synccheck.txt 2>&1; cat gpurun_out/r02_h_synccheck.txt | cut -c1-220
