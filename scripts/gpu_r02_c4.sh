#!/bin/bash
# r02: the fused 2-D acoustic adjoint pass (k_stress2a) on hardware: gradient tests, C4 with and without it, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gradient or pingpong or adjoint or born or boundary" > gpurun_out/pytest_c4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c4.log; tail -3 gpurun_out/pytest_c4.log
for f in 0 1; do
  GPI_FUSE2A=$f timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c4_fuse$f.json 2> gpurun_out/bench_c4_fuse$f.err; echo "c4 fuse=$f rc=$?"; cat gpurun_out/bench_c4_fuse$f.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_fuse1.csv \
    python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c4_fuse1.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stress2a|k_vel2v" -s 40 -c 4 -o gpurun_out/ncu_c4_fuse1 -f \
    python bench.py --workload c4 --nt 40 --nss 16 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c4_full.log 2>&1; echo "ncu full rc=$?"
