#!/usr/bin/env python
"""cold-start probe: the FIRST process on a fresh box; five dense layer steps, every work buffer compared with the last run"""
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
    junk = [torch.full((256 << 20,), float("nan"), device=dev) for _ in range(12)]
    del junk
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
    modes = os.environ.get("MODES", "dense,lowrank,dense,lowrank,dense").split(",")
    for mode in modes:
        junk = [torch.full((256 << 20,), float("nan"), device=dev) for _ in range(12)]
        del junk
        step = LayerStep(At, band, plan, F, F, C, "none", bwd_mode=mode)
        out = step.forward(H, W, U).clone()
        y = step.B1.clone()                       # Y after the forward
        p = step.B2.clone()                       # P
        dH, dW, dU = step.backward(dOut, W, U)
        torch.cuda.synchronize()
        snaps.append({"mode": mode, "out": out, "Y": y, "P": p, "dH": dH.clone(), "B1": step.B1.clone(),
                      "B3": step.B3.clone(), "dW": dW.clone(), "dU": dU.clone()})
        del step
    res = []
    last = {m: [s for s in snaps if s["mode"] == m][-1] for m in set(modes)}
    for i, s in enumerate(snaps):
        r = {"run": i, "mode": s["mode"]}
        for k in ("out", "Y", "P", "dH", "B1", "B3", "dW", "dU"):
            a, b_ = s[k], last[s["mode"]][k]
            same = torch.eq(a, b_) | (torch.isnan(a) & torch.isnan(b_))
            nb = int((~same).sum())
            if nb:
                d = {"n": nb, "maxdiff": float((a - b_).abs().nan_to_num(1e9).max())}
                if a.numel() == T * N * F:
                    rows = torch.nonzero((~same).view(T * N, F).any(1)).flatten()
                    d.update(rows=int(rows.numel()), first=int(rows[0]), last=int(rows[-1]),
                             slices=sorted(set((rows // N).tolist()))[:40])
                r[k] = d
        res.append(r)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
