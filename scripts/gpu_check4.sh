#!/bin/bash
# 2-GPU call: gradient / Born / multi-GPU tests after the merged-pw launches and the weighted slab partition; C4 at small and large batches.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -k "gradient or fwi or born or multigpu or zslab or sharding or boundary" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sel.log
grep -v "^$" gpurun_out/pytest_sel.log | tail -25
for NSS in 4 32; do
timeout 600 python bench.py --workload c4 --nss $NSS --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_c4_nss$NSS.json 2> gpurun_out/bench_c4_nss$NSS.err; echo "c4 nss=$NSS rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c4_nss$NSS.json")); print("nss=$NSS value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "ms", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1))
PY
done
