#!/bin/bash
cd "$(dirname "$0")/.."
for w in 8 12; do
echo "== stem producer warps $w (two m16 tiles in flight)"
YB_STEM_WARPS=$w bash tools/r02_gpu14.sh 2>&1 | grep -E "passed|failed|fused"
done
