#!/bin/bash
# Round 2, programmatic dependent launch of the 2-D chain (GPI_PDL): its own tests, A/B of C2 (1 and 8 resident supersources) and C4, then the
# GPU suite with GPI_PDL=1 (what a default-on build runs), 2-D tests first.  Every step under its own timeout: a dependency bug would hang, not fail.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 150 python -m pytest tests/test_pdl_gpu.py -m gpu -x -q > gpurun_out/pytest_pdl.log 2>&1; echo "pdl tests rc=$?" >> gpurun_out/pytest_pdl.log; tail -2 gpurun_out/pytest_pdl.log
ab() {   # workload nss batch
  for p in 0 1; do
    GPI_PDL=$p timeout 60 python bench.py --workload $1 --nss $2 --shot-batch $3 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 nss $2 pdl $p value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], 'frac', round(r['frac'],3), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
}
{ ab c2 1 1; ab c4 16 16; ab c2 8 8; } 2>&1 | tee gpurun_out/ab_pdl.txt
echo "elapsed $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/ab_pdl.txt
# the 2-D tests first (the attribute is only ever set on 2-D launches), then the 3-D ones (k_post / k_boundary carry the two instructions as no-ops)
K2='not 3d and not c3 and not multigpu and not ragged_z and not pdl'
K3='(3d or c3 or ragged_z) and not multigpu and not pdl'
GPI_PDL=1 timeout 170 python -m pytest tests -m gpu -x -q -k "$K2" > gpurun_out/pytest_r02_pdl1_2d.log 2>&1; echo "2-D suite with GPI_PDL=1 rc=$?" >> gpurun_out/pytest_r02_pdl1_2d.log; tail -3 gpurun_out/pytest_r02_pdl1_2d.log
echo "elapsed $(( $(date +%s) - T0 )) s"
GPI_PDL=1 timeout 200 python -m pytest tests -m gpu -x -q -k "$K3" > gpurun_out/pytest_r02_pdl1_3d.log 2>&1; echo "3-D suite with GPI_PDL=1 rc=$?" >> gpurun_out/pytest_r02_pdl1_3d.log; tail -3 gpurun_out/pytest_r02_pdl1_3d.log
echo "elapsed $(( $(date +%s) - T0 )) s"
