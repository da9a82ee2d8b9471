#!/bin/bash
# Build libyolob200.so (sm_100a) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=yolov3_tensorflow_b200/csrc
OUT=yolov3_tensorflow_b200/libyolob200.so
mkdir -p build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $*"
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -newer "$o" \( -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/*.o -cudart static
echo "built $OUT"
