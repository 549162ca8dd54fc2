#!/bin/bash
# env-only tuning of the C3 time loop (nt = 300): persistent CTA count of the tile kernels, shell placement
mkdir -p gpurun_out; : > gpurun_out/tune_env.txt
run() { echo "== $1" >> gpurun_out/tune_env.txt; env $1 timeout 200 python bench.py --nt 300 --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"value {d['value']:.2f} ms/step {d['ms_per_step']/300*1000:.1f} us/timestep  {r['kernel']} {r['avg_launch_ms']*1000:.1f} us  other {list(r['other'].values())[0]['avg_launch_ms']*1000:.1f} us  both {r['both_kernels_frac']:.3f} clocks {d['clocks']['sm_mhz']}\")" >> gpurun_out/tune_env.txt; }
run "GPI_X=0"
run "GPI_SHELL=0"
run "GPI_TMA3_CTAS=148"
run "GPI_TMA3_CTAS=222"
run "GPI_TMA3_CTAS=592"
run "GPI_SAMPLE_EVERY=0"
run "GPI_X=1"
cat gpurun_out/tune_env.txt
