mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in 0 1; do GPI_SCALAR2D=$v timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    r = d['roofline']; print('SCALAR2D=$v value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), r['kernel'], round(r['avg_launch_ms'],4), 'other', {k: round(v['avg_launch_ms'],4) for k,v in r['other'].items()}, 'both', round(r['both_kernels_frac'],3), 'shots/h', round(d['shots_per_hour']))
"; done
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
