#!/bin/bash
# 2-GPU call: NCCL / z-slab tests, then the C5 z-slab bench and the C4 gradient bench on 2 GPUs.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
grep -v "^$" gpurun_out/pytest_2gpu.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $TR bench.py --gpus 2 --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; echo "c5 rc=$?"
cat gpurun_out/bench_c5_n2.json; tail -3 gpurun_out/bench_c5_n2.err
timeout 600 $TR bench.py --gpus 2 --workload c4 --nss 16 --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; echo "c4 rc=$?"
cat gpurun_out/bench_c4_n2.json; tail -3 gpurun_out/bench_c4_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "c3 rc=$?"
cat gpurun_out/bench_c3_n2.json | cut -c1-500; tail -3 gpurun_out/bench_c3_n2.err
