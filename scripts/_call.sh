mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "full_c3" > gpurun_out/pytest_full.log 2>&1; tail -30 gpurun_out/pytest_full.log
