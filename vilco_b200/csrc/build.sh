#!/bin/bash
# Build libvilco_b200.so in-tree for sm_100a (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${VILCO_NVCC_EXTRA}"
OBJS=""
for f in *.cu; do
  o="build/${f%.cu}.o"
  mkdir -p build
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/vilco_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -shared -o ../libvilco_b200.so $OBJS -lcudart
echo "built $(realpath ../libvilco_b200.so)"
