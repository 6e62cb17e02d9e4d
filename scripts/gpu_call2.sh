#!/bin/bash
# GPU call (2 GPUs): full pytest on one GPU, multi-GPU parity check, 2-GPU bench (default line with extras)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest2.log
for C in 2 3; do
TMGCN_CHECK_CLASSES=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_2gpu_C$C.json 2> gpurun_out/r02_multi_gpu_check_2gpu_C$C.err; echo "check C=$C rc=$?"
tail -c 1500 gpurun_out/r02_multi_gpu_check_2gpu_C$C.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 800 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print(d['ms_per_step'], d['value'], d['halo'], d['parity_multi_gpu'].get('ok'))
print(json.dumps(d['stages_ms_per_rank']))
for k,v in d.get('strong_scaling',{}).items(): print(k, v.get('ms_per_step'), v.get('error'), json.dumps(v.get('stages_ms_per_rank')))
PY
