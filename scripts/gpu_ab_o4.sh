#!/bin/bash
# A/B of library variants / switches on the C3 grid at order 4 (nt = 200): gpu_ab_o4.sh <out> <variant[:ENV=..]> ...
mkdir -p gpurun_out
OUT=gpurun_out/$1; shift
: > $OUT
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=$(echo ${spec#*:} | tr ',' ' ')
  lib=$PWD/geophyinv.jl_b200/variants/lib_$name.so; [ "$name" = default ] && lib=$PWD/geophyinv.jl_b200/libgpifdtd.so
  echo "== $spec" >> $OUT
  env GPI_LIB=$lib $envs timeout 300 python bench.py --order 4 --nt 200 --steps 2 --warmup 2 --no-cpu --no-extra 2>>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"value {d['value']:.2f} us/timestep {d['ms_per_step']/d['config']['time_steps_per_step']*1000:.1f}  {r['kernel']} {r['avg_launch_ms']*1000:.1f} us  other {list(r['other'].values())[0]['avg_launch_ms']*1000:.1f} us  both {r['both_kernels_frac']:.3f}\")" >> $OUT 2>&1
done
cat $OUT
