#!/bin/bash
# First GPU call of round 2 (one GPU, ~12 box-minutes): what was built at the end of round 1 without a GPU gets its hardware run.
#   1. the full GPU test suite (new since the last GPU call: test_pingpong_adjoint_equals_the_copy_path x 4,
#      test_engine_matches_the_stokes_solution x 2, the medium path through gpi_set_medium_fields in every test)
#   2. smoke + the default bench line (C3; e2e now without the host-side numpy passes of update!(pa, medium))
#   3. C4 (FWI gradient) with the save_tp! copy and with GPI_PINGPONG=1: the A/B that decides the default
#   4. ncu launch list of a short C4 gradient in both settings (copy kernel gone, two small boundary launches instead)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02a.log
tail -3 gpurun_out/pytest_r02a.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02a.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02a.log; tail -2 gpurun_out/smoke_r02a.log
timeout 900 python bench.py > gpurun_out/bench_c3_r02a.json 2> gpurun_out/bench_c3_r02a.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3_r02a.json
for pp in 0 1; do
    GPI_PINGPONG=$pp timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c4_pp$pp.json 2> gpurun_out/bench_c4_pp$pp.err
    echo "bench c4 GPI_PINGPONG=$pp rc=$?"; cat gpurun_out/bench_c4_pp$pp.json
    GPI_PINGPONG=$pp timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_pp$pp.csv \
        python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c4_pp$pp.log 2>&1; echo "ncu c4 pp=$pp rc=$?"
done
