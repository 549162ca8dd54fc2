#!/bin/bash
# Builds tuning variants of libgpifdtd.so into geophyinv.jl_b200/variants/ (git-ignored; they travel with gpurun).
# usage: build_variants.sh name1="flags" name2="flags" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p geophyinv.jl_b200/variants
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( GPI_NVCC_EXTRA="$flags" GPI_OUT=../variants/lib_$name.so GPI_TAG=$name bash geophyinv.jl_b200/csrc/build.sh ) &
done
wait
ls -la geophyinv.jl_b200/variants/
