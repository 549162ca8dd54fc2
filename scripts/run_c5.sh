#!/bin/bash
# usage: run_c5.sh N [workload]   -- z-slab bench on N GPUs of this box
N=${1:-2}; WL=${2:-c5}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --workload c5small --steps 1 --warmup 1 > gpurun_out/bench_c5small_n$N.json 2> gpurun_out/bench_c5small_n$N.err; echo "c5small rc=$?"
cat gpurun_out/bench_c5small_n$N.json | cut -c1-600
if [ "$WL" = "c5" ]; then
timeout 1500 $TR bench.py --gpus $N --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 rc=$?"
cat gpurun_out/bench_c5_n$N.json
tail -5 gpurun_out/bench_c5_n$N.err
fi
