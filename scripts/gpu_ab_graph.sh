#!/bin/bash
# CUDA-graph replay of 2-D forward runs: the GPU tests that touch it, then C2 with 1 / 2 / 8 / 16 resident supersources with GPI_GRAPH = 0 and the default
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "graph_replay or acoustic2d or elastic2d or born or golden or order4" 2>&1 | tail -3
for nss in 1 2 8 16; do for g in 0 1; do
  GPI_GRAPH=$g timeout 200 python bench.py --workload c2 --nss $nss --shot-batch $nss --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('c2 nss $nss graph $g value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], 'k_vel2v us', round(r['avg_launch_ms']*1000,1), 'share', round(r['stencil_share_of_step'],3))"
done; done | tee gpurun_out/ab_graph.txt
