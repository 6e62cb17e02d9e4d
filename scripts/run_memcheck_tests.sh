#!/bin/bash
# compute-sanitizer memcheck over the hot path at small sizes (every tensor its own cudaMalloc so an out-of-bounds
# access cannot hide inside the caching allocator's blocks)
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 420 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_smoke.log 2>&1; echo "smoke rc=$?"
grep -c "Invalid\|out of bounds" gpurun_out/r02_memcheck_smoke.log; grep -A12 "Invalid" gpurun_out/r02_memcheck_smoke.log | head -60; tail -4 gpurun_out/r02_memcheck_smoke.log
timeout 700 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "layer_step or fill_variants or mtransform_sparse_matches or stencil or spmm_short or solve_part or sliced or zero_weight" > gpurun_out/r02_memcheck_tests.log 2>&1; echo "tests rc=$?"
grep -c "Invalid\|out of bounds" gpurun_out/r02_memcheck_tests.log; grep -A12 "Invalid" gpurun_out/r02_memcheck_tests.log | head -80; tail -5 gpurun_out/r02_memcheck_tests.log
