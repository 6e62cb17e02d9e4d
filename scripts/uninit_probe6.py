#!/usr/bin/env python
"""dense -> lowrank -> dense under a NaN-poisoned allocator: which SHARED tensor changes between the runs, and which
work buffers differ between the two dense runs"""
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
        if os.environ.get("POISON", "nan") == "none":
            return
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

    def shared(step):
        inc_ptr, perm = step.inc
        return {"H": H, "W": W, "U": U, "dOut": dOut, "At.val": At.val, "At.col": At.col, "At.rowptr": At.rowptr,
                "AtT.val": step.AtT.val, "AtT.col": step.AtT.col, "AtT.rowptr": step.AtT.rowptr, "inc_ptr": inc_ptr,
                "perm": perm, "w_f32": step.w_f32, "src": plan.src, "dst": plan.dst}
    snaps, shareds = [], []
    for mode in ("dense", "lowrank", "dense"):
        poison_pool()
        step = LayerStep(At, band, plan, F, F, C, "none", bwd_mode=mode)
        step.forward(H, W, U)
        dH, dW, dU = step.backward(dOut, W, U)
        torch.cuda.synchronize()
        snaps.append({"dH": dH.clone(), "B1": step.B1.clone(), "B3": step.B3.clone(), "dW": dW.clone(), "dU": dU.clone()})
        shareds.append({k: v.clone() for k, v in shared(step).items()})
        del step
    out = {"shared_changed_after_lowrank": [k for k in shareds[0] if not torch.equal(shareds[0][k], shareds[2][k])],
           "shared_changed_after_dense": [k for k in shareds[0] if not torch.equal(shareds[0][k], shareds[1][k])]}
    for k in ("dH", "B1", "B3", "dW", "dU"):
        a, b_ = snaps[0][k], snaps[2][k]
        same = torch.eq(a, b_)
        d = {"differing": int((~same).sum())}
        if d["differing"] and a.numel() == T * N * F:
            rows = torch.nonzero((~same).view(T * N, F).any(1)).flatten()
            d.update(rows=int(rows.numel()), first=int(rows[0]), last=int(rows[-1]), slices=sorted(set((rows // N).tolist()))[:30],
                     maxdiff=float((a - b_).abs().max()), maxval=float(b_.abs().max()))
        out["dense0_vs_dense2:" + k] = d
    out["lowrank_dH_vs_dense0"] = float((snaps[1]["dH"] - snaps[0]["dH"]).abs().max() / snaps[0]["dH"].abs().max())
    out["lowrank_dH_vs_dense2"] = float((snaps[1]["dH"] - snaps[2]["dH"]).abs().max() / snaps[2]["dH"].abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
