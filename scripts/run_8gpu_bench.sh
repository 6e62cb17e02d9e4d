#!/bin/bash
# GPU call (8 GPUs): the full default bench line (what the driver runs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
grep "\[bench\]" gpurun_out/r02_bench_8gpu.err; grep -i "launch failure\|illegal\|out of memory\|Traceback" gpurun_out/r02_bench_8gpu.err | head -5 | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_8gpu.json'))
print(d['ms_per_step'], d['value'], d['halo']['forward'], d['parity_multi_gpu'].get('ok'), 'dense', d['dense']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d.get('strong_scaling',{}).items(): print(k, v.get('ms_per_step'), v.get('error'), 'dense', (v.get('dense') or {}).get('ms_per_step'))
PY
