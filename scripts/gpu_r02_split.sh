#!/bin/bash
# r02: A/B of the tile kernels with one consumer warp per output group (12 consumer warps per CTA) against the r02a library (4 consumer
# warps), 3-D elastic GPU tests on the new kernels, the default bench line, ncu --set full of the new kernels.
mkdir -p gpurun_out
OLD=$PWD/geophyinv.jl_b200/variants/lib_r02a.so
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"value {d['value']:.2f} us/timestep {d['ms_per_step']/d['config']['time_steps_per_step']*1000:.1f}  {r['kernel']} {r['avg_launch_ms']*1000:.1f} us  other {list(r['other'].values())[0]['avg_launch_ms']*1000:.1f} us  both {r['both_kernels_frac']:.3f} share {r['stencil_share_of_step']:.3f} clocks {d['clocks']['sm_mhz']}\")"; }
timeout 900 python -m pytest tests -m gpu -x -q -k "elastic3d or c3 or tma or adjoint3d or stokes" > gpurun_out/pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_split.log; tail -3 gpurun_out/pytest_split.log
: > gpurun_out/split_ab.txt
for rep in 1 2; do
  echo "== old (r02a) rep $rep" >> gpurun_out/split_ab.txt; GPI_LIB=$OLD timeout 200 python bench.py --nt 400 --steps 2 --warmup 2 --no-cpu --no-extra 2>/dev/null | line >> gpurun_out/split_ab.txt
  echo "== new rep $rep" >> gpurun_out/split_ab.txt; timeout 200 python bench.py --nt 400 --steps 2 --warmup 2 --no-cpu --no-extra 2>/dev/null | line >> gpurun_out/split_ab.txt
done
for c in 148 444; do echo "== new GPI_TMA3_CTAS=$c" >> gpurun_out/split_ab.txt; GPI_TMA3_CTAS=$c timeout 200 python bench.py --nt 400 --steps 2 --warmup 2 --no-cpu --no-extra 2>/dev/null | line >> gpurun_out/split_ab.txt; done
echo "== new GPI_SHELL=0" >> gpurun_out/split_ab.txt; GPI_SHELL=0 timeout 200 python bench.py --nt 400 --steps 2 --warmup 2 --no-cpu --no-extra 2>/dev/null | line >> gpurun_out/split_ab.txt
cat gpurun_out/split_ab.txt
timeout 600 python bench.py --no-extra > gpurun_out/bench_c3_split.json 2> gpurun_out/bench_c3_split.err; echo "bench rc=$?"; cat gpurun_out/bench_c3_split.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step3t -c 4 -o gpurun_out/ncu_c3_split -f \
    python bench.py --nt 6 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_c3_split.log 2>&1; echo "ncu rc=$?"
