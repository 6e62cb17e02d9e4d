#!/usr/bin/env python
"""Multi-GPU parity check of the time-sharded, communication-overlapped layer step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/multi_gpu_check.py

Runs `tmgcn_b200.selfcheck.multi_gpu_parity` (uneven time blocks, NCCL halo and peer-memory fused halo,
general and low-rank backward) and prints one JSON line on rank 0; exit status 1 on a mismatch.
(Not a pytest file: the driver's `-m gpu` box has one GPU; run it with `gpurun --gpus 2`.  `bench.py --gpus N`
runs the same check before timing and reports it as `parity_multi_gpu`.)
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tmgcn_b200 import selfcheck  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    C = int(os.environ.get("TMGCN_CHECK_CLASSES", "2"))
    ok = True
    for _ in range(int(os.environ.get("TMGCN_CHECK_REPEAT", "1"))):
        res = selfcheck.multi_gpu_parity(rank, world, dev, C=C)
        ok = ok and res["ok"]
        if rank == 0:
            print(json.dumps(res), flush=True)
    res["ok"] = ok
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
