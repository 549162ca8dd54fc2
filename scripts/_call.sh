mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "3d or c3 or elastic or acou3" 2>&1 | tail -6
TUNE_NT=200 timeout 900 bash scripts/tune.sh
