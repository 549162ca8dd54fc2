#!/bin/bash
# Round 2, one GPU: the full GPU test suite at HEAD, smoke, the default bench line (C3), C2 and C4 lines, ncu launch lists of C3 / C4 and
# --set full captures of the kernels that changed this round (k_step3t<0|1>, k_stress2a, k_vel2v<0,1>, k_post).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02_final.log
tail -3 gpurun_out/pytest_r02_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_final.log; tail -2 gpurun_out/smoke_r02_final.log
timeout 900 python bench.py > gpurun_out/bench_c3_r02_final.json 2> gpurun_out/bench_c3_r02_final.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3_r02_final.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c4_r02_final.json 2>/dev/null; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4_r02_final.json
timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_c2_r02_final.json 2>/dev/null; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2_r02_final.json
timeout 600 python bench.py --order 4 --nt 300 --steps 2 --warmup 2 --no-cpu --no-extra > gpurun_out/bench_c3_o4_r02_final.json 2>/dev/null; echo "bench o4 rc=$?"; cat gpurun_out/bench_c3_o4_r02_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3_r02_final.csv \
    python bench.py --workload c3 --nt 50 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l_c3.log 2>&1; echo "ncu list c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_r02_final.csv \
    python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_l_c4.log 2>&1; echo "ncu list c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step3t -s 8 -c 2 -o gpurun_out/ncu_c3_r02_final -f \
    python bench.py --nt 8 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_f_c3.log 2>&1; echo "ncu full c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stress2a|k_post" -s 6 -c 3 -o gpurun_out/ncu_c4_r02_final -f \
    python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_f_c4.log 2>&1; echo "ncu full c4 rc=$?"
