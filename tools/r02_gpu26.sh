#!/bin/bash
# which tensor fails test_train_step_matches_oracle[flags1]?  plain vs split-precision stem; then the whole suite and the A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "YB_STEM_SPLIT=0" "YB_STEM_SPLIT=1" "YB_STEM_TRAIN=cuda"; do
  echo "--- [$cfg]"
  env $cfg timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider --tb=short -x -k "test_train_step_matches_oracle and flags1" 2>&1 | grep -E "^E  |passed|failed|  0 |  1 |  2 |  3 " | cut -c1-400 | head -14
done > gpurun_out/r02_o_stem_parity.txt 2>&1; cat gpurun_out/r02_o_stem_parity.txt
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -12 > gpurun_out/r02_o_tests.log; tail -6 gpurun_out/r02_o_tests.log | cut -c1-400
echo "=== train A/B"
timeout 900 python tools/train_ab.py 32 416 10 -- "" "YB_STEM_SPLIT=0" > gpurun_out/r02_o_train_ab.txt 2>&1; cat gpurun_out/r02_o_train_ab.txt | cut -c1-200
