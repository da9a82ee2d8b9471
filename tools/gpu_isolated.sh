#!/bin/bash
# Run every `-m gpu` test of the given files in its own process (a trapping kernel kills the
# CUDA context of its process only).  Output: gpurun_out/isolated.log + one summary line per test.
mkdir -p gpurun_out
LOG=gpurun_out/isolated.log
: > $LOG
ids=$(python -m pytest "$@" --collect-only -q -m gpu 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  out=$(timeout 180 python -m pytest "$id" -q -x 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; else
    fail=$((fail+1)); echo "FAIL($rc) $id"
    echo "==================== $id (rc=$rc)" >> $LOG
    echo "$out" | grep -v "^  " | tail -120 >> $LOG
  fi
done
echo "isolated: $pass passed, $fail failed"
