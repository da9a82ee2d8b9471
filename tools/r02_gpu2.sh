#!/bin/bash
# round-2 GPU session 2: previously failing tests with full tracebacks, in-kernel timelines of the 1x1 layers, ncu --set full of two 1x1 layers
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 900 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider -k "train_step_matches_oracle or forward_parity or frozen_bn or optimizer_zoo or multi_scale or checkpoint" --tb=short 2>&1 | tail -150 > gpurun_out/r02_2_tests.log
tail -12 gpurun_out/r02_2_tests.log
echo "=== traces"
T=gpurun_out/r02_2_traces.txt; : > $T
tr() { timeout 120 python tools/conv_trace.py "$@" >> $T 2>&1; }
tr 64 52 52 256 128 1 1
YB_CONV_DBG=7 tr 64 52 52 256 128 1 1
YB_CONV_EPI=reg tr 64 52 52 256 128 1 1
YB_CONV_MODE=1cta tr 64 26 26 512 256 1 1
tr 64 104 104 128 64 1 1
tr 64 52 52 128 256 3 1 res
YB_CONV_MODE=1cta tr 64 52 52 128 256 3 1 res
tr 64 208 208 32 64 3 1 res
cat $T | cut -c1-260
echo "=== ncu"
for spec in "64 52 52 256 128 1 1" "64 26 26 512 256 1 1"; do
  tag=$(echo $spec | tr ' ' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 2 -c 1 -f -o gpurun_out/r02_2_ncu_$tag python tools/conv_probe.py $spec 1 > gpurun_out/r02_2_ncu_$tag.log 2>&1
  tail -2 gpurun_out/r02_2_ncu_$tag.log
done
ls -la gpurun_out | tail -8
