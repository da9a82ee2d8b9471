#!/bin/bash
# round-2 GPU session 9: fine-grained stamps inside one epilogue chunk
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=gpurun_out/r02_9_traces.txt; : > $T
trace() { timeout 120 python tools/conv_trace.py "$@" 2>&1 | head -8 | cut -c1-330 >> $T; }
YB_CONV_EG=1 trace 64 52 52 256 128 1 1
YB_CONV_EG=1 YB_CONV_DBG=7 trace 64 52 52 256 128 1 1
trace 64 52 52 256 128 1 1
cat $T
