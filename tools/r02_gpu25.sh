#!/bin/bash
# wgrad on a side stream beside dgrad: tests, training step A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1200 python -m pytest tests/test_gpu_path.py tests/test_gpu_train_ops.py tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -8 > gpurun_out/r02_n_tests.log; tail -4 gpurun_out/r02_n_tests.log | cut -c1-300
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_WGRAD_STREAM=0" "" > gpurun_out/r02_n_train_ab.txt 2>&1; cat gpurun_out/r02_n_train_ab.txt | cut -c1-200
