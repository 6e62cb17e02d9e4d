#!/usr/bin/env python
"""Multi-GPU parity check of the time-sharded, communication-overlapped layer step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/multi_gpu_check.py

Every rank builds the same seeded graph, takes its time block (with the sparse halo from its
predecessor over NCCL), runs LayerStep.forward/backward with a ShardComm, and the gathered
results are compared on rank 0 with the single-shard run of the whole tensor.
(Not a pytest file: the driver's `-m gpu` box has one GPU; run it with `gpurun --gpus 2`.)
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg  # noqa: E402
from tmgcn_b200 import ops, sharding, synth  # noqa: E402
from tmgcn_b200.layer_step import LayerStep  # noqa: E402


MODE_ACT = os.environ.get("TMGCN_CHECK_ACT", "relu")   # "none" exercises the low-rank backward + its skinny halo


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    N, F, C, b = 20000, 128, 3, 5
    T = 12 * world
    idx, val = synth.synth_coo(N, T, 60000, 0.85, seed=11, device="cpu")
    M = tg.create_matrix_M(T, b)
    band = tg.Band(M)
    g = torch.Generator().manual_seed(3)
    H = torch.rand(T, N, F, generator=g)
    W = (torch.randn(F, F, generator=g) / F ** 0.5).to(dev)
    U = torch.randn(2 * F, C, generator=g).to(dev)
    full_A = tg.SliceCSR.from_coo(idx, val, T, N)
    full_At = ops.mtransform_sparse(full_A, band)
    E = 40000
    edges = synth.synth_edges(full_At, E, seed=5)          # same on every rank (same seed, same tensor)
    dOut = torch.randn(E, C, generator=g).to(dev)

    t0, t1 = sharding.shard_bounds(T, world, rank)
    Tl = t1 - t0
    halo = min(b - 1, t0) if rank > 0 else 0
    sel = (idx[0] >= t0) & (idx[0] < t1)
    own_idx = idx[:, sel].clone()
    own_idx[0] -= t0
    A_own = tg.SliceCSR.from_coo(own_idx, val[sel], Tl, N)
    A_in = sharding.exchange_sparse_halo(A_own, b - 1, rank, world)
    assert A_in.T == Tl + halo
    At = ops.mtransform_sparse(A_in, band, t0, t1, halo)
    esel = (edges[0] >= t0) & (edges[0] < t1)
    plan = tg.EdgePlan(edges[:, esel], N, t_offset=t0)
    step = LayerStep(At, band, plan, F, F, C, MODE_ACT, t0, t1, halo)
    Hl = torch.zeros(halo + Tl, N, F, device=dev)
    Hl[halo:] = H[t0:t1].to(dev)                           # the halo slices arrive over NCCL
    comm = sharding.ShardComm(b - 1, rank, world, dev)
    for _ in range(2):                                     # twice: buffers and events are re-used
        out = step.forward(Hl, W, U, comm).clone()
        dH, dW, dU = step.backward(dOut[esel].contiguous(), W, U, comm)
        torch.cuda.synchronize()
    assert torch.equal(Hl[:halo].cpu(), H[t0 - halo:t0]), "forward halo content"
    # the same forward with the halo exchange fused into the stencil over NVLink peer memory
    peer = sharding.PeerHalo(Tl, N, F, b - 1, rank, world, dev)
    peer.H.copy_(H[t0:t1].to(dev))
    for _ in range(2):
        out_p = step.forward(peer.H, W, U, comm, peer).clone()
        torch.cuda.synchronize()
    assert torch.equal(out_p, out), "peer-memory halo: forward differs from the NCCL halo path"
    dH_p, dW_p, dU_p = step.backward(dOut[esel].contiguous(), W, U, comm)
    torch.cuda.synchronize()
    assert torch.equal(dW_p, dW) and torch.equal(dU_p, dU)
    print(f"rank {rank}: peer-memory fused halo == NCCL halo (bit-identical logits)", flush=True)

    # reference: the whole tensor on one GPU
    ok = True
    ref = LayerStep(full_At, band, tg.EdgePlan(edges, N), F, F, C, MODE_ACT, bwd_mode="dense")
    out_r = ref.forward(H.to(dev), W, U).clone()
    dH_r, dW_r, dU_r = ref.backward(dOut, W, U)

    def rel(a, b_):
        return ((a.double() - b_.double()).abs().max() / b_.double().abs().max()).item()
    errs = {"out": rel(out, out_r[esel]), "dH": rel(dH[halo:], dH_r[t0:t1]), "dW": rel(dW, dW_r), "dU": rel(dU, dU_r)}
    ok = errs["out"] <= 1e-5 and max(errs["dH"], errs["dW"], errs["dU"]) <= 1e-4
    print(f"rank {rank}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + (" OK" if ok else " FAIL"), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
