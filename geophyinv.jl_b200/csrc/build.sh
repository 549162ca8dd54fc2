#!/bin/bash
# Builds libgpifdtd.so (sm_100a) in-tree next to the Python mirror.  nvcc cross-compiles without a GPU.
#   GPI_NVCC_EXTRA : extra nvcc flags (e.g. -DGPI_VEC_THREADS=128 -DGPI_VEC_MINBLOCKS=2)
#   GPI_OUT        : output file (default ../libgpifdtd.so); tuning variants are loaded with GPI_LIB=<path>
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${GPI_OUT:-../libgpifdtd.so}
LOG=build${GPI_TAG:+_$GPI_TAG}.log
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC -shared -Xptxas -v ${GPI_NVCC_EXTRA} \
      -o $OUT engine.cu -ldl 2> $LOG || { cat $LOG; exit 1; }
grep -E "error|warning" $LOG | grep -v "Function properties" | grep -v nonnull | head -20 || true
echo "built $OUT"
