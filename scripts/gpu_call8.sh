#!/bin/bash
# GPU call (8 GPUs): multi-GPU parity check, default bench (weak c5shard + strong c5cut/c4 + parity), c4 variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_8gpu.json 2> gpurun_out/r02_multi_gpu_check_8gpu.err; echo "check rc=$?"
tail -c 1200 gpurun_out/r02_multi_gpu_check_8gpu.json
timeout 900 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 400 gpurun_out/r02_bench_8gpu.err
i=0
for V in "0" "148" "592"; do
  i=$((i+1))
  TMGCN_BOUNDARY_CTAS=$V timeout 400 $TR --master-port $((29520+i)) bench.py --gpus 8 --preset c4 --no-extras --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c4_ctas$V.json 2> gpurun_out/r02_bench_8gpu_c4_ctas$V.err; echo "c4 ctas=$V rc=$?"
done
timeout 400 $TR --master-port 29530 bench.py --gpus 8 --preset c4 --no-extras --halo nccl --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c4_nccl.json 2> gpurun_out/r02_bench_8gpu_c4_nccl.err; echo "c4 nccl rc=$?"
timeout 400 $TR --master-port 29531 bench.py --gpus 8 --preset c4 --no-extras --bwd dense --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_c4_dense.json 2> gpurun_out/r02_bench_8gpu_c4_dense.err; echo "c4 dense rc=$?"
python - <<'PY'
import json,glob
d=json.load(open('gpurun_out/r02_bench_8gpu.json'))
print('weak', d['ms_per_step'], d['value'], d['halo']['forward'], d['parity_multi_gpu'].get('ok'))
for r in d['stages_ms_per_rank']: print(json.dumps(r))
for k,v in d.get('strong_scaling',{}).items():
    print(k, v.get('ms_per_step'), v.get('error'))
    for r in (v.get('stages_ms_per_rank') or []): print('  ', json.dumps(r))
for f in sorted(glob.glob('gpurun_out/r02_bench_8gpu_c4_*.json')):
    try:
        x=json.load(open(f)); print(f, x['ms_per_step'], x['halo']['forward'], x['backward']['mode'])
    except Exception as e: print(f, 'ERR', e)
PY
