#!/bin/bash
# usage: gpu_ngpu.sh N -- N-GPU call: (N >= 4: 4-rank tests), C5 z-slabs, C2 (8 supersources per GPU), C4 (32/N... supersources per GPU)
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
if [ "$N" = "4" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s -k "four" > gpurun_out/pytest_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${N}gpu.log
grep -v "^$" gpurun_out/pytest_${N}gpu.log | tail -8
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 rc=$?"
cat gpurun_out/bench_c5_n$N.json | cut -c1-1800; tail -3 gpurun_out/bench_c5_n$N.err
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "c2 rc=$?"
cat gpurun_out/bench_c2_n$N.json | cut -c1-900; tail -3 gpurun_out/bench_c2_n$N.err
timeout 600 $TR bench.py --gpus $N --workload c4 --nss $((32 / N)) --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"
cat gpurun_out/bench_c4_n$N.json | cut -c1-1500; tail -3 gpurun_out/bench_c4_n$N.err
