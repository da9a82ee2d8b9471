#!/bin/bash
# BN streaming-kernel rework + fused finalize + multi-tensor dgrad-weight repack: tests, micro-benchmark, step A/B; synccheck experiments
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -15 > gpurun_out/r02_g_tests.log; tail -6 gpurun_out/r02_g_tests.log | cut -c1-300
echo "=== bn probe"
timeout 600 python tools/bn_probe.py 7 > gpurun_out/r02_g_bn_probe.md 2>&1; cat gpurun_out/r02_g_bn_probe.md | cut -c1-200
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_BN_CPT=8" "YB_BN_FIN=0" "YB_PACK_MT=0" "YB_BN_CPT=8,YB_BN_FIN=0,YB_PACK_MT=0" "" > gpurun_out/r02_g_train_ab.txt 2>&1; cat gpurun_out/r02_g_train_ab.txt | cut -c1-200
echo "=== synccheck experiments on the flagged test"
for cfg in "" "YB_CONV_DBG=64" "YB_CONV_EG=1"; do
  echo "--- [$cfg]"
  env $cfg timeout 300 compute-sanitizer --tool synccheck --print-limit 1 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "conv1x1_slices_residual and 1cta and not eg and not reg" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Barrier error|by thread|located|Device Frame.*kernel" | head -8
done > gpurun_out/r02_g_synccheck.txt 2>&1; cat gpurun_out/r02_g_synccheck.txt | cut -c1-220
