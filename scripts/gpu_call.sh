#!/bin/bash
# One gpurun call: GPU parity tests, the C3 bench, tuning variants, the 2-D bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
TUNE_NT=${TUNE_NT:-300} timeout 900 bash scripts/tune.sh
if [ -n "$RUN_C2" ]; then
timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
fi
