#!/bin/bash
# wave-aware tile width for the pair kernel: tests, inference and training step A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -8 > gpurun_out/r02_m_tests.log; tail -4 gpurun_out/r02_m_tests.log | cut -c1-300
echo "=== infer A/B"
timeout 900 python tools/infer_ab.py 64 416 10 -- "" "YB_CONV_WAVE=0" "" > gpurun_out/r02_m_infer_ab.txt 2>&1; cat gpurun_out/r02_m_infer_ab.txt | cut -c1-200
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_CONV_WAVE=0" > gpurun_out/r02_m_train_ab.txt 2>&1; cat gpurun_out/r02_m_train_ab.txt | cut -c1-200
