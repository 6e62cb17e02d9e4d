#!/usr/bin/env python
"""stage-by-stage version of uninit_probe.py for the dense backward: which stage's output depends on what the
freshly allocated work buffers happened to contain"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg
from tmgcn_b200 import ops, synth, _lib
from tmgcn_b200.ops import _p, _stream


def main():
    T, N, F, C, b = 99, 20000, 128, 2, 5
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
    band = tg.Band(tg.create_matrix_M(T, b))
    At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
    AtT = At.transpose()
    g = torch.Generator().manual_seed(3)
    dP = torch.randn(T, N, F, generator=g).to(dev)
    w = band.device_weights(0, T, torch.float32)
    out = {}
    ref = {}
    for fill in (0.0, float("nan"), 1e30):
        dHt = torch.full((T, N, F), fill, device=dev)
        dH = torch.full((T, N, F), fill, device=dev)
        _lib.check(lib.tmgcn_spmm_fwd(_p(AtT.rowptr), _p(AtT.col), _p(AtT.val), _p(dP), _p(dHt), T, N, F, 0, _stream()))
        a = dHt.clone()
        _lib.check(lib.tmgcn_mtransform_dense_bwd_range(_p(dHt), _p(dH), T, 0, N * F, _p(w), band.b, 0, T, -1, _stream()))
        torch.cuda.synchronize()
        bq = dH.clone()
        key = str(fill)
        if not ref:
            ref = {"spmm": a, "stencil": bq}
        out[key] = {"spmm_equal": bool(torch.equal(a, ref["spmm"])), "stencil_equal": bool(torch.equal(bq, ref["stencil"])),
                    "stencil_nan": int(torch.isnan(bq).sum()), "spmm_nan": int(torch.isnan(a).sum()),
                    "stencil_maxdiff": float((bq - ref["stencil"]).abs().nan_to_num(1e9).max())}
        # the plain (non-ranged) entry point too
        dH2 = torch.full((T, N, F), fill, device=dev)
        _lib.check(lib.tmgcn_mtransform_dense_bwd(_p(dHt), _p(dH2), T, 0, N * F, _p(w), band.b, _stream()))
        out[key]["plain_equal"] = bool(torch.equal(dH2, ref["stencil"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
