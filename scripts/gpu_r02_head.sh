#!/bin/bash
# Round 2, last call: HEAD (programmatic dependent launch on by default, NVTX switch, merged Stokes run) -- the GPU tests that changed or that
# the changes touch, smoke, the headline C3 line (no extra legs) and the C4 / C2 lines.  Sized for the 3.7 GPU-minutes the round had left.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 100 python -m pytest tests -m gpu -x -q -k "stokes or pdl or graph_replay or fwi_gradient_acoustic2d or c1_acoustic or multishot" > gpurun_out/pytest_r02_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02_head.log; tail -2 gpurun_out/pytest_r02_head.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke_r02_head.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_head.log; tail -2 gpurun_out/smoke_r02_head.log
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 60 python bench.py --workload c4 --no-cpu > gpurun_out/bench_c4_r02_head.json 2> gpurun_out/bench_c4_r02_head.err; echo "c4 rc=$?"; cut -c1-200 gpurun_out/bench_c4_r02_head.json
GPI_NVTX=1 timeout 40 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2_r02_head.json 2> gpurun_out/bench_c2_r02_head.err; echo "c2 rc=$?"; cut -c1-200 gpurun_out/bench_c2_r02_head.json
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 90 python bench.py --no-extra > gpurun_out/bench_c3_r02_head.json 2> gpurun_out/bench_c3_r02_head.err; echo "c3 rc=$?"; cut -c1-200 gpurun_out/bench_c3_r02_head.json
echo "elapsed $(( $(date +%s) - T0 )) s"
