#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tune_shell.txt
for lib in geophyinv.jl_b200/libgpifdtd.so geophyinv.jl_b200/variants/lib_shell6.so geophyinv.jl_b200/variants/lib_shell8.so; do
  echo "== $lib" >> gpurun_out/tune_shell.txt
  GPI_LIB=$PWD/$lib timeout 200 python bench.py --nt 300 --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"value {d['value']:.2f} {d['ms_per_step']/300*1000:.1f} us/timestep  {r['kernel']} {r['avg_launch_ms']*1000:.1f} us  other {list(r['other'].values())[0]['avg_launch_ms']*1000:.1f} us  both {r['both_kernels_frac']:.3f}\")" >> gpurun_out/tune_shell.txt
done
cat gpurun_out/tune_shell.txt
