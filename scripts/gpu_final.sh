#!/bin/bash
# Round-end check on one GPU: full GPU test suite, smoke, the default bench line (C3), its ncu launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -3 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
cat gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --workload c3 --nt 50 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches_final.log 2>&1; echo "ncu rc=$?"
