#!/bin/bash
# One `ncu --set full` capture per stencil kernel family that the C3 profile (prof_c3_tma) does not cover.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
run() {   # name regex skip count case
  timeout 400 $NCU -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1 python scripts/ncu_cases.py $5 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
}
run acou2d      'k_(vel|stress)2v'            20 2 acou2d
run elastic2d   'k_(vel|stress)2v'            20 2 elastic2d
run acou3d      'k_(vel|stress)3v'            10 2 acou3d
run el3d_o4     'k_(vel4|stress4|dirichlet4)' 10 5 elastic3d_o4
run acou2d_o4   'k_(vel4|stress4)'            20 2 acou2d_o4
run grad2d      'k_(grad2d|boundary)'         60 3 gradient2d
run grad2d_adj  'k_(vel|stress)2v'           100 2 gradient2d
run grad3d_el   'k_(grad3d_el|boundary)'      20 3 gradient3d_el
ls -la gpurun_out/*.ncu-rep
