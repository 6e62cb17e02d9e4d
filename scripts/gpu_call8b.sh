#!/bin/bash
# GPU call (8 GPUs): final multi-GPU numbers of the round
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_8gpu.json 2> gpurun_out/r02_multi_gpu_check_8gpu.err; echo "check rc=$?"
tail -c 500 gpurun_out/r02_multi_gpu_check_8gpu.json; grep -i "Traceback" -A12 gpurun_out/r02_multi_gpu_check_8gpu.err | tail -30
timeout 900 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
grep -i "Traceback" -A12 gpurun_out/r02_bench_8gpu.err | tail -30
TMGCN_PEER_SIGNALS=0 timeout 400 $TR --master-port 29521 bench.py --gpus 8 --preset c4 --no-extras --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c4_barriers.json 2> gpurun_out/r02_bench_8gpu_c4_barriers.err; echo "c4 barriers rc=$?"
timeout 400 $TR --master-port 29522 bench.py --gpus 8 --preset c4 --no-extras --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c4_signals.json 2> gpurun_out/r02_bench_8gpu_c4_signals.err; echo "c4 signals rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json,glob
for f in ('gpurun_out/r02_bench_8gpu.json','gpurun_out/r02_bench_4gpu.json'):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['value'], d['halo']['forward'], d['parity_multi_gpu'].get('ok'), 'dense', d['dense']['ms_per_step'], 'host', d['host_enqueue_ms_per_step'])
        for r in d['stages_ms_per_rank']: print('  ', json.dumps({k:v for k,v in r.items() if v>0.3}))
        for r in d['dense']['stages_ms_per_rank']: print('  D', json.dumps({k:v for k,v in r.items() if v>0.5}))
        for k,v in d.get('strong_scaling',{}).items():
            print(k, v.get('ms_per_step'), v.get('error'), 'dense', (v.get('dense') or {}).get('ms_per_step'))
            for r in (v.get('stages_ms_per_rank') or []): print('    ', json.dumps({k2:v2 for k2,v2 in r.items() if v2>0.05}))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/r02_bench_8gpu_c4_*.json')):
    try:
        x=json.load(open(f)); print(f, x['ms_per_step'], x['halo']['forward'], x['host_enqueue_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
