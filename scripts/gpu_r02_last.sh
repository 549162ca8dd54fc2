#!/bin/bash
# Round 2, last single-GPU call at HEAD: full GPU suite, smoke, the default bench line, --set full capture of k_stress2a as shipped
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02_last.log; tail -3 gpurun_out/pytest_r02_last.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_r02_last.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_last.log; tail -2 gpurun_out/smoke_r02_last.log
timeout 300 python bench.py > gpurun_out/bench_c3_r02_last.json 2> gpurun_out/bench_c3_r02_last.err; echo "bench rc=$?"; cat gpurun_out/bench_c3_r02_last.json | cut -c1-900
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_stress2a" -s 4 -c 1 -o gpurun_out/ncu_c4_stress2a_last -f \
    python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f_c4_last.log 2>&1; echo "ncu rc=$?"
