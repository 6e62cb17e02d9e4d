#!/bin/bash
# GPU call (1 GPU): ncu launch list + DRAM traffic of the default bench, one --set full capture of every kernel
mkdir -p gpurun_out
KREG='regex:spmm_|stencil_|gemm_|merge_|readout_|factor_|row_class|scan_|rowptr_|transpose_|sgemm|dw_|reduce_partials|act_|flat_ids|csr_|solve_|gather_'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREG" -c 700 \
    --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 \
    > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k "$KREG" -c 80 -o gpurun_out/r02_prof \
    python bench.py --no-extras --no-cpu-baseline --slices 8 --steps 1 --warmup 3 \
    > gpurun_out/r02_bench_under_ncu_full.json 2> gpurun_out/r02_bench_under_ncu_full.err; echo "ncu full rc=$?"
ls -la gpurun_out/r02_prof* gpurun_out/r02_launches.csv
