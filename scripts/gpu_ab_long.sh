#!/bin/bash
# A/B of library variants on the full C3 bench step (2000 time steps: the board reaches its power cap), same box: gpu_ab_long.sh <out> <variant[:ENV=..]> ...
mkdir -p gpurun_out
OUT=gpurun_out/$1; shift
: > $OUT
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=$(echo ${spec#*:} | tr ',' ' ')
  lib=$PWD/geophyinv.jl_b200/variants/lib_$name.so; [ "$name" = default ] && lib=$PWD/geophyinv.jl_b200/libgpifdtd.so
  echo "== $spec" >> $OUT
  env GPI_LIB=$lib $envs timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-extra 2>>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"value {d['value']:.2f} us/timestep {d['ms_per_step']/d['config']['time_steps_per_step']*1000:.1f}  {r['kernel']} {r['avg_launch_ms']*1000:.1f} us  other {list(r['other'].values())[0]['avg_launch_ms']*1000:.1f} us  both {r['both_kernels_frac']:.3f} share {r['stencil_share_of_step']:.3f} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}\")" >> $OUT 2>&1
done
cat $OUT
