mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "3d or c3 or elastic or acou3" 2>&1 | tail -4
TUNE_NT=200 timeout 900 bash scripts/tune.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step3t' -s 20 -c 2 -f -o gpurun_out/prof_t3 python bench.py --workload c3 --nt 30 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full_t3.log 2>&1
