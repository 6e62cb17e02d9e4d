#!/usr/bin/env python
"""is the per-slice transpose (built once, cached) sensitive to what the allocator hands it?"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg
from tmgcn_b200 import ops, synth


def main():
    T, N, b = 99, 20000, 5
    dev = torch.device("cuda", 0)
    idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
    band = tg.Band(tg.create_matrix_M(T, b))
    At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
    torch.cuda.synchronize()
    out = {}
    ref = None
    for tag, fill in (("clean", None), ("nan", float("nan")), ("big", 1e30), ("nan2", float("nan"))):
        if fill is not None:
            junk = [torch.full((256 << 20,), fill, device=dev) for _ in range(12)]
            del junk
        A2 = tg.SliceCSR(At.T, At.N, At.rowptr, At.col, At.val)      # fresh object: no cached transpose
        tr = A2.transpose()
        torch.cuda.synchronize()
        snap = (tr.rowptr.clone(), tr.col.clone(), tr.val.clone())
        if ref is None:
            ref = snap
            # ground truth: transpose of the COO by sorting (t, j, i)
            i3, v3 = At.to_coo()
            key = (i3[0] * N + i3[2]) * N + i3[1]
            order = torch.argsort(key)
            out["clean_matches_sort"] = bool(torch.equal(snap[1].long(), i3[1][order]) and torch.equal(snap[2], v3[order]))
        out[tag] = {"rowptr": bool(torch.equal(snap[0], ref[0])), "col": bool(torch.equal(snap[1], ref[1])),
                    "val": bool(torch.equal(snap[2], ref[2])), "n_col_diff": int((snap[1] != ref[1]).sum()),
                    "n_val_diff": int((snap[2] != ref[2]).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
