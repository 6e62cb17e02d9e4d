#!/usr/bin/env python
"""Poison-the-allocator probe: fill most of the caching allocator with NaN, free it, then run the layer step in both
backward formulations (single GPU) and compare -- a kernel that reads a workspace it never wrote shows up as NaN /
garbage, and as a run-to-run difference."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg  # noqa: E402
from tmgcn_b200 import ops, synth  # noqa: E402
from tmgcn_b200.layer_step import LayerStep  # noqa: E402


def rel(a, b):
    d = b.double().abs().max().item()
    return (a.double() - b.double()).abs().max().item() / (d if d > 0 else 1.0)


def main():
    poison = os.environ.get("POISON", "nan")
    T, N, F, C, b = int(os.environ.get("PT", "99")), 20000, 128, 2, 5
    dev = torch.device("cuda", 0)
    idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
    band = tg.Band(tg.create_matrix_M(T, b))

    def poison_pool():
        if poison == "none":
            return
        junk = [torch.full((256 << 20,), float("nan") if poison == "nan" else 1e30, device=dev) for _ in range(12)]
        del junk
    poison_pool()
    full_At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
    g = torch.Generator().manual_seed(3)
    H = torch.rand(T, N, F, generator=g).to(dev)
    W = (torch.randn(F, F, generator=g) / F ** 0.5).to(dev)
    U = torch.randn(2 * F, C, generator=g).to(dev)
    E = 2 * N
    edges = synth.synth_edges(full_At, E, seed=5)
    dOut = torch.randn(E, C, generator=g).to(dev)
    plan = tg.EdgePlan(edges, N, T=T)
    out = {}
    res = {}
    for mode in ("dense", "lowrank", "dense", "lowrank"):
        poison_pool()
        step = LayerStep(full_At, band, plan, F, F, C, "none", bwd_mode=mode)
        o = step.forward(H, W, U).clone()
        dH, dW, dU = step.backward(dOut, W, U)
        torch.cuda.synchronize()
        key = mode + ("2" if mode in res else "")
        res[key] = (o, dH.clone(), dW.clone(), dU.clone())
        del step
    for a, bname in (("lowrank", "dense"), ("dense2", "dense"), ("lowrank2", "lowrank")):
        out[a + "_vs_" + bname] = {n: rel(x, y) for n, x, y in zip(("out", "dH", "dW", "dU"), res[a], res[bname])}
    out["finite"] = all(bool(torch.isfinite(t).all()) for v in res.values() for t in v)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
