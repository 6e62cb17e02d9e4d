#!/bin/bash
# GPU call (2 GPUs): validate the reworked communication + new kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest3.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_2gpu_b.json 2> gpurun_out/r02_multi_gpu_check_2gpu_b.err; echo "check rc=$?"
tail -c 700 gpurun_out/r02_multi_gpu_check_2gpu_b.json; grep -i "error\|Traceback" -A5 gpurun_out/r02_multi_gpu_check_2gpu_b.err | tail -20
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench2 rc=$?"
grep -i "error\|Traceback" -A8 gpurun_out/r02_bench_2gpu_b.err | tail -30
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_b.json'))
print(d['ms_per_step'], d['value'], d['parity_multi_gpu'].get('ok'), d['host_enqueue_ms_per_step'], 'dense', d['dense']['ms_per_step'])
print(json.dumps(d['stages_ms_per_rank'])); print(json.dumps(d['dense']['stages_ms_per_rank']))
for k,v in d.get('strong_scaling',{}).items(): print(k, v.get('ms_per_step'), v.get('host_enqueue_ms_per_step'), v.get('error'), (v.get('dense') or {}).get('ms_per_step'))
PY
