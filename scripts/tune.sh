#!/bin/bash
# Runs the C3 bench (short) once per library variant in geophyinv.jl_b200/variants/ and prints kernel times.
mkdir -p gpurun_out
: > gpurun_out/tune.jsonl
for lib in geophyinv.jl_b200/libgpifdtd.so geophyinv.jl_b200/variants/lib*.so; do
  echo "== $lib" >> gpurun_out/tune.jsonl
  GPI_LIB=$PWD/$lib timeout 300 python bench.py --nt ${TUNE_NT:-300} --steps 1 --warmup 1 --no-cpu 2>> gpurun_out/tune.err >> gpurun_out/tune.jsonl
done
python - <<'PY'
import json
name=None
for l in open('gpurun_out/tune.jsonl'):
    if l.startswith('=='): name=l.strip(); continue
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print(f"{name:55s} value {d['value']:.2f} Gcell/s  {r['kernel']} {r['avg_launch_ms']:.4f} ms frac {r['frac']:.3f}  other {list(r['other'].values())[0]['avg_launch_ms']:.4f} ms  both {r['both_kernels_frac']:.3f}")
PY
