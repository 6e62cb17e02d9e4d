#!/bin/bash
# GPU call (1 GPU): ncu launch list + DRAM traffic of the default bench step, one --set full capture of every kernel.
# The .ncu-rep stays on the box (it is larger than what gpurun copies back); CSV exports come home.
mkdir -p gpurun_out
KREG='regex:spmm_|stencil_|gemm_|merge_|count_union|fill_from_union|fill_union|readout_|factor_|row_class|scan_|rowptr_|transpose_|sgemm|dw_|reduce_partials|act_|flat_ids|csr_|solve_|gather_'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREG" -c 400 \
    --csv --log-file gpurun_out/r02_launches.csv python bench.py --lean --steps 2 --warmup 3 \
    > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k "$KREG" -c 70 -o /tmp/r02_prof \
    python bench.py --lean --slices 8 --steps 1 --warmup 3 \
    > gpurun_out/r02_bench_under_ncu_full.json 2> gpurun_out/r02_bench_under_ncu_full.err; echo "ncu full rc=$?"
ncu -i /tmp/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_prof_raw.csv 2> /dev/null; echo "raw export rc=$?"
ncu -i /tmp/r02_prof.ncu-rep --page source --csv -k regex:merge_rows > gpurun_out/r02_prof_source_merge.csv 2> /dev/null; echo "source export rc=$?"
ls -la gpurun_out/ /tmp/r02_prof.ncu-rep
