#!/bin/bash
# round-2 GPU session 7: ablation of the TMA epilogue chunk (what takes ~1100 cycles?) + eval / preprocess GPU tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 600 python -m pytest tests/test_eval_utils.py tests/test_gpu_preprocess.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-250
echo "=== traces"
T=gpurun_out/r02_7_traces.txt; : > $T
trace() { timeout 120 python tools/conv_trace.py "$@" 2>&1 | head -7 >> $T; }
YB_CONV_DBG=7 trace 64 52 52 256 128 1 1
YB_CONV_DBG=23 trace 64 52 52 256 128 1 1
YB_CONV_DBG=55 trace 64 52 52 256 128 1 1
YB_CONV_DBG=39 trace 64 52 52 256 128 1 1
cut -c1-250 $T
