#!/usr/bin/env python
"""which work buffer of the dense backward differs between a run on NaN-poisoned fresh buffers and a run on reused ones"""
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
    for fill in (float("nan"), 0.0):
        step = LayerStep(At, band, plan, F, F, C, "none", bwd_mode="dense")
        for buf in (step.B1, step.B2, step.B3, step.out, step.dW, step.dU):
            buf.fill_(fill)
        step.du_ws.view(torch.float32).fill_(fill) if step.du_ws.numel() % 4 == 0 else None
        step.dw_ws.view(torch.float32).fill_(fill) if step.dw_ws.numel() % 4 == 0 else None
        step.forward(H, W, U)
        step.backward(dOut, W, U)
        torch.cuda.synchronize()
        snaps.append({k: getattr(step, k).clone() for k in ("B1", "B2", "B3", "dW", "dU")})
        del step
    out = {}
    for k in ("B1", "B2", "B3", "dW", "dU"):
        a, b_ = snaps[0][k], snaps[1][k]
        same = torch.eq(a, b_) | (torch.isnan(a) & torch.isnan(b_))
        n_bad = int((~same).sum())
        out[k] = {"differing": n_bad, "nan": int(torch.isnan(a).sum())}
        if n_bad and a.numel() == T * N * F:
            rows = torch.nonzero((~same).view(T * N, F).any(1)).flatten()
            out[k].update(rows=int(rows.numel()), first=int(rows[0]), last=int(rows[-1]),
                          mod128=sorted(set((rows % 128).tolist()))[:12], slices=sorted(set((rows // N).tolist()))[:12])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
