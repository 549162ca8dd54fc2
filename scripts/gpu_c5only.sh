#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5_n${N}_w.json 2> gpurun_out/bench_c5_n${N}_w.err; echo "c5 rc=$?"
cat gpurun_out/bench_c5_n${N}_w.json | cut -c1-1900; tail -3 gpurun_out/bench_c5_n${N}_w.err
