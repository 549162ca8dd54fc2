#!/bin/bash
# Experiment: 2-D shot batch size vs L2 residency (C2 grid, 16 supersources), and order-4 throughput for reference.
mkdir -p gpurun_out
for B in 1 2 4 8 16; do
  timeout 300 python bench.py --workload c2 --nss 16 --shot-batch $B --nt 1000 --steps 2 --warmup 3 --no-cpu > gpurun_out/c2_b$B.json 2> gpurun_out/c2_b$B.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c2_b$B.json")); r=d.get("roofline") or {}
    print("B=$B value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "kernels ms", r.get("avg_launch_ms"), (r.get("other") or {}))
except Exception as e:
    print("B=$B failed", e)
PY
done
timeout 300 python bench.py --workload c2 --order 4 --nt 1000 --steps 2 --warmup 3 --no-cpu > gpurun_out/c2_o4.json 2> gpurun_out/c2_o4.err; cat gpurun_out/c2_o4.json | cut -c1-400
timeout 300 python bench.py --workload c3 --order 4 --nt 100 --steps 2 --warmup 3 --no-cpu > gpurun_out/c3_o4.json 2> gpurun_out/c3_o4.err; cat gpurun_out/c3_o4.json | cut -c1-400; tail -2 gpurun_out/c3_o4.err
