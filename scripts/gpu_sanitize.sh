#!/bin/bash
# compute-sanitizer passes over small runs of every kernel family + the new GPU tests.  Outputs under gpurun_out/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
grep -E "ok,|ERROR SUMMARY|Invalid|rc=|DONE" gpurun_out/memcheck.log | tail -30
timeout 600 $CS --tool racecheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_cases.py elastic3d > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
grep -E "ok,|RACECHECK SUMMARY|hazard|rc=|DONE" gpurun_out/racecheck.log | tail -12
timeout 600 $CS --tool initcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_cases.py elastic3d acou2d > gpurun_out/initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/initcheck.log
grep -E "ok,|ERROR SUMMARY|Uninitialized|rc=|DONE" gpurun_out/initcheck.log | tail -12
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "simultaneous" 2>&1 | tail -5
