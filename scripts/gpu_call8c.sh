#!/bin/bash
# diagnostic (8 GPUs): which leg of the default workload fails, with and without neighbour signals
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 300 $TR --master-port 29513 bench.py --gpus 8 --no-extras --steps 6 --warmup 3 > gpurun_out/r02_diag_signals.json 2> gpurun_out/r02_diag_signals.err; echo "signals rc=$?"
grep "\[bench\]" gpurun_out/r02_diag_signals.err; grep -i "launch failure\|illegal\|out of memory\|NCCL WARN" gpurun_out/r02_diag_signals.err | head -5 | cut -c1-300
TMGCN_PEER_SIGNALS=0 NCCL_DEBUG=WARN timeout 300 $TR --master-port 29514 bench.py --gpus 8 --no-extras --steps 6 --warmup 3 > gpurun_out/r02_diag_barriers.json 2> gpurun_out/r02_diag_barriers.err; echo "barriers rc=$?"
grep "\[bench\]" gpurun_out/r02_diag_barriers.err; grep -i "launch failure\|illegal\|out of memory\|NCCL WARN" gpurun_out/r02_diag_barriers.err | head -5 | cut -c1-300
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv | head -10
