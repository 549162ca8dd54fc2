#!/bin/bash
# r02, 2 GPUs: multi-GPU tests at HEAD (slab windows now pick the 96-cell tiles), then the slab shape C5 has on 8 GPUs (75 owned planes,
# z pitch 96) on a 150 x 1106 x 1106 grid over 2 slabs: register-staged kernels vs 96-cell TMA tiles vs 128-cell TMA tiles
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_2gpu_slab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu_slab.log
grep -v "^$" gpurun_out/pytest_2gpu_slab.log | tail -6
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for spec in "GPI_TMA3=0" "GPI_TMA3=1" "GPI_TMA3=2,GPI_TMA3_ZC=128"; do
  envs=$(echo $spec | tr ',' ' ')
  echo "== $spec"
  env $envs timeout 600 $TR bench.py --gpus 2 --workload c5 --c5-ni 68,1024,1024 --nt 100 --steps 2 --warmup 1 2>>gpurun_out/c5shape.err | tee -a gpurun_out/c5shape.raw | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(f\"value {d['value']:.1f} ms/timestep {d['ms_per_time_step']:.3f} kernel {r['kernel']} {r['avg_launch_ms']:.3f} ms other {list(r['other'].values())[0]['avg_launch_ms']:.3f} ms both {r['both_kernels_frac']:.3f} whole {r['whole_step_frac']:.3f} exch/step {d['exchange_ms_per_time_step']:.3f} ms share {d['exchange_share']:.3f}\")" | tee -a gpurun_out/c5shape.txt
done
