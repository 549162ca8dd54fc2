#!/bin/bash
# Round 2, 2-GPU call: the multi-GPU pytest cases at HEAD, then the default bench line under torchrun with its C4 / C5 legs
# (parity flags, exchange share), then the reference arm under torchrun (thread count must be the box's, not OMP_NUM_THREADS=1).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_2gpu_r02.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu_r02.log
grep -v "^$" gpurun_out/pytest_2gpu_r02.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1500 $TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err; echo "bench rc=$?"
cat gpurun_out/bench_n2_r02.json; tail -5 gpurun_out/bench_n2_r02.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2_r02.json 2> gpurun_out/bench_ref_n2_r02.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_n2_r02.json; tail -3 gpurun_out/bench_ref_n2_r02.err
