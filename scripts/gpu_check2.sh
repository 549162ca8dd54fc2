#!/bin/bash
# Runs on the B200 box under gpurun: GPU parity tests, smoke, the C3/C2/C4 bench lines, the ncu launch list and one
# --set full capture of the TMA-pipelined 3-D stencil kernels (k_step3t).  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-2} --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?" >> gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?" >> gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?" >> gpurun_out/bench_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --workload c3 --nt 60 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step3t' -s 20 -c 4 -f -o gpurun_out/prof_c3_tma \
    python bench.py --workload c3 --nt 40 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
cat gpurun_out/bench_c2.json; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err
