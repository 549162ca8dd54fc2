#!/bin/bash
# A/B of library variants on a short C4 gradient (16 supersources, nt = 600): gpu_ab_c4.sh <out-file> <variant> ...
mkdir -p gpurun_out
OUT=gpurun_out/$1; shift
: > $OUT
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=$(echo ${spec#*:} | tr ',' ' ')
  lib=$PWD/geophyinv.jl_b200/variants/lib_$name.so; [ "$name" = default ] && lib=$PWD/geophyinv.jl_b200/libgpifdtd.so
  echo "== $spec" >> $OUT
  env GPI_LIB=$lib $envs timeout 200 python bench.py --workload c4 --nt 600 --nss 16 --steps 2 --warmup 1 --no-cpu 2>>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(f\"value {d['value']:.2f} e2e {d['e2e']['value']:.2f} ms/step {d['ms_per_step']:.1f} launches {d['gpu_launches']} frac {d['roofline']['frac']:.3f}\")" >> $OUT 2>&1
done
cat $OUT
