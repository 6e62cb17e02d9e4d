#!/bin/bash
# GPU call: tests + default bench + probes (1 GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_a.err
python scripts/spmm_chunk_probe.py > gpurun_out/r02_spmm_chunk_probe.json 2> gpurun_out/r02_spmm_chunk_probe.err; echo "probe rc=$?"
python scripts/measure_tf32_peak.py > gpurun_out/r02_tf32_peak.json 2>&1; echo "tf32 rc=$?"
cat gpurun_out/r02_spmm_chunk_probe.json gpurun_out/r02_tf32_peak.json
