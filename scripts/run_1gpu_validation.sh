#!/bin/bash
# GPU call (1 GPU): full GPU test suite + the default bench line + smoke
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest5.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['layer_hbm_frac'], 'dense', d['dense']['ms_per_step'])
print('mtransform', d['mtransform_sparse']['seconds'], d['mtransform_sparse']['hbm_frac'])
print('e2e', d['e2e']['ms_per_step'], 'fresh', json.dumps(d.get('e2e_fresh_inputs'))[:600])
print('same', d['extra'], 'cpu', d['cpu_baseline']['value'])
print({k:v.get('ms_per_step') for k,v in d['strong_scaling'].items()}, d.get('gemm_tensor_pipe',{}).get('executed_frac_of_sustained_peak'))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_ref.json
