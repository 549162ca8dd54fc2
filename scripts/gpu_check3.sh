#!/bin/bash
# GPU parity tests (all, including order 4 and the batched boundary store) and the C4 / C2 bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?" >> gpurun_out/bench_c4.err
cat gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err
