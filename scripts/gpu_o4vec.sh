#!/bin/bash
mkdir -p gpurun_out
GPI_O4VEC=1 timeout 600 python -m pytest tests/test_order4.py tests/test_adjoint3d.py tests/test_golden.py -m gpu -q -k "order4 or order or o4" > gpurun_out/pytest_o4vec.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_o4vec.log
tail -4 gpurun_out/pytest_o4vec.log
for V in 0 1; do
GPI_O4VEC=$V timeout 200 python bench.py --order 4 --nt 100 --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('C3 order 4 O4VEC=$V value %.2f  %s %.1f us  other %.1f us  both %.3f' % (d['value'], r['kernel'], r['avg_launch_ms']*1000, list(r['other'].values())[0]['avg_launch_ms']*1000, r['both_kernels_frac']))"
done
GPI_O4VEC=1 timeout 200 python bench.py --workload c2 --order 4 --nt 1000 --steps 2 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 order 4 O4VEC=1 value %.2f' % d['value'])"
