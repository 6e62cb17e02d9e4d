#!/usr/bin/env python
"""target for compute-sanitizer: sparse transform + dense and low-rank layer steps at the self-check's shape"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg
from tmgcn_b200 import ops, synth
from tmgcn_b200.layer_step import LayerStep

T, N, F, C, b = int(os.environ.get("PT", "27")), int(os.environ.get("PN", "20000")), 128, 2, 5
dev = torch.device("cuda", 0)
idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
band = tg.Band(tg.create_matrix_M(T, b))
At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
# the union-list kernels' overflow paths: a hub row (its block misses the staging pool -> fill_union_overflow) ...
hub = torch.stack([torch.randint(0, T, (4000,)), torch.full((4000,), 7), torch.randint(0, N, (4000,))])
Ch = torch.sparse_coo_tensor(torch.cat([idx, hub], 1), torch.cat([val, torch.rand(4000, dtype=torch.float64)]),
                             (T, N, N)).coalesce()
for dt in (torch.float32, torch.float64):
    o = ops.mtransform_sparse(tg.SliceCSR.from_coo(Ch._indices(), Ch._values(), T, N, dtype=dt), band)
    print("hub", dt, o.nnz, float(o.val.double().sum()))
del Ch, hub, o
g = torch.Generator().manual_seed(3)
H = torch.rand(T, N, F, generator=g).to(dev)
W = (torch.randn(F, F, generator=g) / F ** 0.5).to(dev)
U = torch.randn(2 * F, C, generator=g).to(dev)
E = 2 * N
edges = synth.synth_edges(At, E, seed=5)
dOut = torch.randn(E, C, generator=g).to(dev)
plan = tg.EdgePlan(edges, N, T=T)
for mode in ("dense", "lowrank"):
    step = LayerStep(At, band, plan, F, F, C, "none", bwd_mode=mode)
    step.forward(H, W, U)
    dH, dW, dU = step.backward(dOut, W, U)
    torch.cuda.synchronize()
    print(mode, float(dH.abs().sum()), float(dW.abs().sum()), float(dU.abs().sum()))
    del step
