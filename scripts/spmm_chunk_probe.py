#!/usr/bin/env python
"""Probe for the "L2-blocked SpMM" idea: time the facewise SpMM kernel at feature widths
F = 4 ... 128 on the benchmark graph (N = 2M, ~40M stored entries per transformed slice).  At F <= 8 one
slice of the operand (N * 4F bytes <= 64 MB) is L2-resident, so a pass at that width IS one feature chunk of
the blocked variant: 128/F such passes would replace the one F = 128 pass.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tmgcn_b200 as tg  # noqa: E402
from tmgcn_b200 import ops, synth  # noqa: E402


def main():
    N, m, T, b = 2_000_000, 10_000_000, 12, 10
    tg._lib.load(build_if_missing=False)
    A = synth.synth_csr(N, T, m, 0.9, seed=20261017, t_start=40)
    band = tg.Band(tg.create_matrix_M(T, b))
    At = ops.mtransform_sparse(A, band)
    del A
    # keep only the 3 full-window slices (t >= b-1)
    nnz_t = At.slice_nnz().tolist()
    res = {"N": N, "nnz_per_slice": nnz_t[-1], "passes": {}}
    T_use = 3
    r0 = (T - T_use) * N
    base = int(At.rowptr[r0].item())
    sub = tg.SliceCSR(T_use, N, (At.rowptr[r0:] - base).contiguous(), At.col[base:].contiguous(),
                      At.val[base:].contiguous())
    nnz = sub.nnz
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for F in (4, 8, 16, 32, 64, 128):
        x = torch.rand(T_use, N, F, device="cuda")
        y = torch.empty_like(x)
        for _ in range(2):
            ops.spmm_raw(sub, x, 0, y)
        ev0.record()
        for _ in range(5):
            ops.spmm_raw(sub, x, 0, y)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5 / T_use
        alg = (8.0 * nnz + 4.0 * F * nnz + 8.0 * N * F * T_use) / T_use     # operand gathered from L2 or DRAM
        res["passes"][str(F)] = {"ms_per_slice": ms, "passes_for_F128": 128 // F, "blocked_ms_per_slice": ms * (128 // F),
                                 "gathered_GBps": 4.0 * F * (nnz / T_use) / ms / 1e6, "operand_slice_MB": N * 4 * F / 1e6}
        del x, y
    print(json.dumps(res))


if __name__ == "__main__":
    main()
