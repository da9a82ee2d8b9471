#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck + synccheck over a slice of the conv
# tests (1-CTA / pair / halo kernels, both epilogues, residual, statistics), the fused detection step and one training step.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SEL='conv1x1_gemm_basic or conv1x1_slices_residual or conv3x3_s1_im2col or conv3x3_s2 or conv3x3_bf16_and_stats or conv1x1_head_fp32_255 or conv1x1_upsample_into_concat or halo'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool: conv tests"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "($SEL) and (1cta or 2cta) and not eg1 and not mc" > gpurun_out/r02_sanitizer_${tool}_conv.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/r02_sanitizer_${tool}_conv.log | cut -c1-200
done
for tool in memcheck racecheck; do
  echo "=== $tool: detect + train step"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider -x -k "detect_fused_equals_unfused_pipeline and 64-96 or train_step_frozen_bn_fixed_bar and bf16 or nms_edge_cases" > gpurun_out/r02_sanitizer_${tool}_path.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/r02_sanitizer_${tool}_path.log | cut -c1-200
done
grep -h "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_sanitizer_*.log
