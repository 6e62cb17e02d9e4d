"""CPU model of the union-list sparse M-transform (count record + lane-per-entry fill),
held to the oracle's func_MProduct.  It pins the identities the CUDA kernels rely on, with the kernels' own index
arithmetic (slots, windows, iteration-major record, row-major walk in chunks of 32 lanes):

  * the union pattern of a (4 output slices x 32 rows) task, with a hit mask per entry, determines every output;
  * the value of source slot k for a union entry with bit k sits at
        first entry of the block in that slice + number of earlier union entries with bit k
    (rows are stored back to back, in the order of the union list), so no merge is needed to find it;
  * output tt's entries are the union entries whose mask meets tt's non-zero-weight slots, at base + rank.

Pure numpy / Python loops on small cases: test infrastructure, like oracle/."""
import numpy as np
import pytest
import torch

import oracle
from tmgcn_b200 import synth

TT = 4


def _group_setup(band_w, b, Bt, t0, T_out, halo):
    NS = Bt - 1 + TT
    w = np.zeros((TT, Bt))
    nz = [0] * TT
    used = [False] * NS
    for tt in range(TT):
        for j in range(Bt):                       # j-th slot of output tt's window: slot tt + j, lag Bt-1-j
            lag = Bt - 1 - j
            sl = halo + t0 + tt - lag
            wl = band_w[t0 + tt, lag] if (lag < b and t0 + tt < T_out and sl >= 0) else 0.0
            w[tt, j] = wl
            if wl != 0.0:
                nz[tt] |= 1 << (tt + j)
                used[tt + j] = True
    return w, nz, used


def union_transform(in_rowptr, in_col, in_val, band_w, b, Bt, T_out, halo, N):
    NS = Bt - 1 + TT
    nblk, n_groups = (N + 31) // 32, (T_out + TT - 1) // TT
    counts = np.zeros(T_out * N, dtype=np.int64)
    records = {}
    for task in range(n_groups * nblk):           # ---- count pass: merge once, record {column, hit mask}
        blk, g = divmod(task, n_groups)
        t0 = g * TT
        _, nz, used = _group_setup(band_w, b, Bt, t0, T_out, halo)
        rec, ulen = {}, [0] * 32
        for lane in range(32):
            i = blk * 32 + lane
            if i >= N:
                continue
            lists = [list(in_col[in_rowptr[(halo + t0 - (Bt - 1) + k) * N + i]:
                                 in_rowptr[(halo + t0 - (Bt - 1) + k) * N + i + 1]]) if used[k] else []
                     for k in range(NS)]
            pos, q, cnt = [0] * NS, 0, [0] * TT
            while True:
                cur = [lists[k][pos[k]] if pos[k] < len(lists[k]) else 2 ** 31 - 1 for k in range(NS)]
                m = min(cur)
                if m == 2 ** 31 - 1:
                    break
                hits = 0
                for k in range(NS):
                    if cur[k] == m:
                        hits |= 1 << k
                        pos[k] += 1
                rec[(q, lane)] = (m, hits)        # iteration-major record
                q += 1
                for tt in range(TT):
                    cnt[tt] += 1 if hits & nz[tt] else 0
            ulen[lane] = q
            for tt in range(TT):
                if t0 + tt < T_out:
                    counts[(t0 + tt) * N + i] = cnt[tt]
        records[task] = (ulen, rec)
    rowptr = np.concatenate([[0], np.cumsum(counts)])
    out_col = np.full(rowptr[-1], -1, dtype=np.int64)
    out_val = np.full(rowptr[-1], np.nan)
    for task in range(n_groups * nblk):           # ---- fill pass: one lane per union entry, no merge
        blk, g = divmod(task, n_groups)
        t0, r0 = g * TT, blk * 32
        ulen, rec = records[task]
        roff = np.concatenate([[0], np.cumsum(ulen)])
        packed = [rec[(j, r)] for r in range(32) for j in range(ulen[r])]      # the row-major walk
        assert len(packed) == roff[-1]
        w, nz, used = _group_setup(band_w, b, Bt, t0, T_out, halo)
        sb = [int(in_rowptr[(halo + t0 - (Bt - 1) + k) * N + r0]) if used[k] else 0 for k in range(NS)]
        ob = [int(rowptr[(t0 + tt) * N + r0]) if t0 + tt < T_out else 0 for tt in range(TT)]
        for e0 in range(0, len(packed), 32):
            chunk = packed[e0:e0 + 32]
            vd = np.zeros((len(chunk), NS))
            for k in range(NS):
                bal = [bool(mk & (1 << k)) for _, mk in chunk]
                for lane, bit in enumerate(bal):
                    if bit:
                        vd[lane, k] = float(in_val[sb[k] + sum(bal[:lane])])   # running base + ballot rank
                sb[k] += sum(bal)
            for tt in range(TT):
                bal = [bool(mk & nz[tt]) for _, mk in chunk]
                for lane, bit in enumerate(bal):
                    if bit:
                        acc = 0.0
                        for j in range(Bt):                                    # ascending source-slice order
                            acc = acc + w[tt, j] * vd[lane, tt + j]
                        p = ob[tt] + sum(bal[:lane])
                        out_col[p], out_val[p] = chunk[lane][0], acc
                ob[tt] += sum(bal)
    return counts, out_col, out_val


@pytest.mark.parametrize("T,N,m,rho,b,Bt,cut,zero_w", [
    (9, 70, 300, 0.8, 3, 4, None, False),         # template width above the band width
    (10, 45, 200, 0.5, 5, 6, None, True),         # explicit zeros inside the band
    (11, 64, 250, 0.9, 4, 4, 5, False),           # a rank's block with a halo; N a multiple of 32
    (7, 33, 150, 0.0, 2, 2, 1, False),            # no persistence, one row in the last block, halo shorter than b-1
])
def test_union_list_formulation_equals_func_mproduct(T, N, m, rho, b, Bt, cut, zero_w):
    idx, val = synth.synth_coo(N, T, m, rho, seed=T + N)
    M = oracle.create_matrix_M(T, b).double()
    if zero_w:
        M[2, 1] = 0.0
        M[T - 1, T - 2] = 0.0
    ref_idx, ref_val = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy(), no_diag=b)
    idx, val32 = idx.numpy(), val.numpy().astype(np.float32)
    rowptr = np.zeros(T * N + 1, dtype=np.int64)
    np.add.at(rowptr, idx[0] * N + idx[1] + 1, 1)
    rowptr = np.cumsum(rowptr)
    bw = np.zeros((T, b))
    for lag in range(b):
        bw[lag:, lag] = torch.diagonal(M, -lag).numpy()
    t_lo, halo = (0, 0) if cut is None else (cut, min(b - 1, cut))
    T_out, s0 = T - t_lo, t_lo - halo
    base = rowptr[s0 * N]
    counts, out_col, out_val = union_transform(rowptr[s0 * N:] - base, idx[2][base:], val32[base:], bw[t_lo:], b, Bt,
                                               T_out, halo, N)
    sel = ref_idx[0] >= t_lo
    rows = np.repeat(np.arange(T_out * N), counts)
    got = np.stack([rows // N + t_lo, rows % N, out_col])
    assert np.array_equal(got, ref_idx[:, sel])
    np.testing.assert_allclose(out_val, ref_val[sel], rtol=2e-7, atol=0)
