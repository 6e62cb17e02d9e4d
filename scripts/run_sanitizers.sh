#!/bin/bash
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
PT=99 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_target.py > gpurun_out/r02_memcheck_T99.log 2>&1; echo "memcheck rc=$?"
grep -A14 "Invalid\|Error:" gpurun_out/r02_memcheck_T99.log | head -50; tail -3 gpurun_out/r02_memcheck_T99.log
PT=12 PN=6000 timeout 500 compute-sanitizer --tool initcheck --print-limit 20 python scripts/sanitize_target.py > gpurun_out/r02_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -B2 -A14 "Uninitialized" gpurun_out/r02_initcheck.log | head -90; tail -3 gpurun_out/r02_initcheck.log
PT=8 PN=3000 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_target.py > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -B2 -A10 "hazard\|Race" gpurun_out/r02_racecheck.log | head -60; tail -3 gpurun_out/r02_racecheck.log
