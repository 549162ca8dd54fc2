#!/bin/bash
# r02, 4 GPUs: the multi-GPU tests (middle ranks exchange with both neighbours through the pipelined path), then C5 over 4 z-slabs
# with the pipelined and the serial exchange, then the default bench line under torchrun with its C4 / C5 legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_4gpu_r02.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_4gpu_r02.log
grep -v "^$" gpurun_out/pytest_4gpu_r02.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
for spec in "GPI_SLAB_PIPE=1" "GPI_SLAB_PIPE=0"; do
  echo "== C5 over 4 slabs, $spec"
  env $spec timeout 900 $TR bench.py --gpus 4 --workload c5 --nt 100 --steps 2 --warmup 1 2>>gpurun_out/c5_n4.err | tee -a gpurun_out/bench_c5_n4_r02.raw | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(f\"value {d['value']:.1f} ms/timestep {d['ms_per_time_step']:.3f} {r['kernel']} {r['avg_launch_ms']:.3f} ms other {list(r['other'].values())[0]['avg_launch_ms']:.3f} ms both {r['both_kernels_frac']:.3f} whole {r['whole_step_frac']:.3f} exch/step {d['exchange_ms_per_time_step']:.3f} ms\")"
done
timeout 1500 $TR bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_n4_r02.json 2> gpurun_out/bench_n4_r02.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n4_r02.json').read().splitlines() if l.startswith('{')][-1])
print('C3 weak x4', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
for k,v in d.get('extra',{}).items():
    print(k, {q: (round(v[q],3) if isinstance(v[q],float) else v[q]) for q in v if q in ('value','parity_ok','slab_parity_ok','bit_identical','exchange_share','ms_per_time_step','n_gpus')})
PY
