#!/bin/bash
# Builds libgpifdtd.so (sm_100a) in-tree next to the Python mirror.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libgpifdtd.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC -shared -Xptxas -v ${GPI_NVCC_EXTRA} \
      -o $OUT engine.cu -ldl 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Function properties" | head -20 || true
echo "built $OUT"
