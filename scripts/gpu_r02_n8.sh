#!/bin/bash
# r02, 8 GPUs: the default bench line under torchrun with its C4 / C5 legs (bounded: the whole call has < 4 minutes of box time left)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 190 $TR bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_n8_r02.json 2> gpurun_out/bench_n8_r02.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8_r02.json').read().splitlines() if l.startswith('{')][-1])
    print('C3 weak x8', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
    for k,v in d.get('extra',{}).items():
        print(k, {q: (round(v[q],3) if isinstance(v[q],float) else v[q]) for q in v if q in ('value','parity_ok','slab_parity_ok','bit_identical','exchange_share','ms_per_time_step','n_gpus','failed','leg_wall_s')})
except Exception as e:
    print('no line', e)
PY
tail -3 gpurun_out/bench_n8_r02.err
