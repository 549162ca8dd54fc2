#!/bin/bash
# 2 GPUs: A/B of library variants on the slab shape C5 has on 8 GPUs (150 x 1106 x 1106 grid over 2 slabs of 75 planes): gpu_ab_slab.sh <out> <variant[:ENV=..]> ...
mkdir -p gpurun_out
OUT=gpurun_out/$1; shift
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=$(echo ${spec#*:} | tr ',' ' ')
  lib=$PWD/geophyinv.jl_b200/variants/lib_$name.so; [ "$name" = default ] && lib=$PWD/geophyinv.jl_b200/libgpifdtd.so
  echo "== $spec" >> $OUT
  env GPI_LIB=$lib $envs timeout 300 $TR bench.py --gpus 2 --workload c5 --c5-ni ${C5NI:-68,1024,1024} --nt 100 --steps 2 --warmup 1 2>>gpurun_out/ab_slab.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(f\"value {d['value']:.1f} ms/timestep {d['ms_per_time_step']:.3f} {r['kernel']} {r['avg_launch_ms']:.3f} ms other {list(r['other'].values())[0]['avg_launch_ms']:.3f} ms both {r['both_kernels_frac']:.3f} whole {r['whole_step_frac']:.3f} exch/step {d['exchange_ms_per_time_step']:.3f} ms share {d['exchange_share']:.3f}\")" >> $OUT 2>&1
done
cat $OUT
