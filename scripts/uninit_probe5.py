#!/usr/bin/env python
"""probe 1 again (NaN-poisoned allocator, dense backward twice) with the work buffers compared"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg
from tmgcn_b200 import ops, synth
from tmgcn_b200.layer_step import LayerStep


def main():
    T, N, F, C, b = 99, 20000, 128, 2, 5
    dev = torch.device("cuda", 0)
    idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
    band = tg.Band(tg.create_matrix_M(T, b))

    def poison_pool():
        junk = [torch.full((256 << 20,), float("nan"), device=dev) for _ in range(12)]
        del junk
    poison_pool()
    At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
    g = torch.Generator().manual_seed(3)
    H = torch.rand(T, N, F, generator=g).to(dev)
    W = (torch.randn(F, F, generator=g) / F ** 0.5).to(dev)
    U = torch.randn(2 * F, C, generator=g).to(dev)
    E = 2 * N
    edges = synth.synth_edges(At, E, seed=5)
    dOut = torch.randn(E, C, generator=g).to(dev)
    plan = tg.EdgePlan(edges, N, T=T)
    snaps = []
    for it in range(3):
        poison_pool()
        step = LayerStep(At, band, plan, F, F, C, "none", bwd_mode="dense")
        step.forward(H, W, U)
        step.backward(dOut, W, U)
        torch.cuda.synchronize()
        snaps.append({k: getattr(step, k).clone() for k in ("B1", "B2", "B3", "dW", "dU")})
        snaps[-1]["atT_val"] = step.AtT.val.clone()
        snaps[-1]["atT_col"] = step.AtT.col.clone()
        del step
    out = {}
    for j in (1, 2):
        for k in snaps[0]:
            a, b_ = snaps[0][k], snaps[j][k]
            same = torch.eq(a, b_) | (torch.isnan(a.float()) & torch.isnan(b_.float()))
            n_bad = int((~same).sum())
            d = {"differing": n_bad}
            if n_bad and a.numel() == T * N * F:
                rows = torch.nonzero((~same).view(T * N, F).any(1)).flatten()
                d.update(rows=int(rows.numel()), first=int(rows[0]), last=int(rows[-1]),
                         slices=sorted(set((rows // N).tolist()))[:20],
                         maxdiff=float((a - b_).abs().max()), maxval=float(b_.abs().max()))
            out[f"run0_vs_run{j}:{k}"] = d
    print(json.dumps(out))


if __name__ == "__main__":
    main()
