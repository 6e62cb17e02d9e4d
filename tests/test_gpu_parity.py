"""Parity of the CUDA path (through the C ABI) against the oracle and the golden
fixtures.  Bars: sparse indices bit-exact; embeddings / logits
||d||_inf / ||ref||_inf <= 1e-5; gradients <= 1e-4; identical argmax wherever the
top-2 logit margin exceeds 1e-4 * ||out||_inf."""
import numpy as np
import pytest
import torch

import oracle
from conftest import MPRODUCT_CASES

pytestmark = pytest.mark.gpu

TOL_OUT = 1e-5
TOL_GRAD = 1e-4


@pytest.fixture(scope="module")
def tg():
    import tmgcn_b200
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (the product has no CPU fallback)")
    tmgcn_b200._lib.load(build_if_missing=False)   # the shipped .so must be there: no fallback
    return tmgcn_b200


def relerr(got, ref):
    got = got.detach().double().cpu() if torch.is_tensor(got) else torch.as_tensor(got).double()
    ref = ref.detach().double().cpu() if torch.is_tensor(ref) else torch.as_tensor(ref).double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    denom = ref.abs().max().item()
    return (got - ref).abs().max().item() / (denom if denom > 0 else 1.0)


def assert_same_argmax(got, ref):
    got, ref = got.detach().cpu().double(), torch.as_tensor(ref).double()
    top2 = torch.topk(ref, 2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref.abs().max()
    assert torch.equal(got.argmax(1)[sure], ref.argmax(1)[sure])


# --------------------------------------------------------------------------
# integer plumbing
# --------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 5, 4096, 4097, 1_000_003])
def test_exclusive_scan(tg, n):
    from tmgcn_b200 import ops
    g = torch.Generator().manual_seed(n)
    c = torch.randint(0, 50, (n,), generator=g, dtype=torch.int64)
    out = ops.exclusive_scan(c.cuda()).cpu()
    ref = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(c, 0)])
    assert torch.equal(out, ref)


def test_csr_from_coo_ragged(tg):
    # empty first/last slices, empty rows, a row with many entries
    T, N = 5, 7
    idx = torch.tensor([[1, 1, 1, 3, 3], [0, 6, 6, 2, 2], [3, 0, 5, 1, 2]])
    val = torch.arange(5, dtype=torch.float64)
    csr = tg.SliceCSR.from_coo(idx, val, T, N)
    rp = csr.rowptr.cpu()
    assert rp[0] == 0 and rp[-1] == 5 and rp.numel() == T * N + 1
    counts = (rp[1:] - rp[:-1])
    assert counts[1 * N + 0] == 1 and counts[1 * N + 6] == 2 and counts[3 * N + 2] == 2 and counts.sum() == 5
    idx2, val2 = csr.to_coo()
    assert torch.equal(idx2.cpu(), idx) and torch.equal(val2.cpu().double(), val)
    empty = tg.SliceCSR.from_coo(torch.zeros(3, 0, dtype=torch.int64), torch.zeros(0), 2, 3)
    assert empty.nnz == 0 and torch.equal(empty.rowptr.cpu(), torch.zeros(7, dtype=torch.int64))


# --------------------------------------------------------------------------
# (a) sparse M-transform: bit-exact indices
# --------------------------------------------------------------------------
@pytest.mark.parametrize("name", MPRODUCT_CASES)
def test_func_mproduct_golden(tg, golden_mproduct, name):
    g = golden_mproduct
    T, N, _ = (int(x) for x in g[name + "_shape"])
    C = torch.sparse_coo_tensor(torch.from_numpy(g[name + "_in_idx"]), torch.from_numpy(g[name + "_in_val"]),
                                (T, N, N)).coalesce()
    out = tg.func_MProduct(C, torch.from_numpy(g[name + "_M"]), no_diag=int(g[name + "_b"]))
    assert out.is_coalesced() and out.dtype == torch.float64
    assert out._indices().dtype == torch.int64
    assert torch.equal(out._indices().cpu(), torch.from_numpy(g[name + "_out_idx"]))
    np.testing.assert_allclose(out._values().cpu().numpy(), g[name + "_out_val"], rtol=1e-13, atol=0)


def test_func_mproduct_chess(tg, golden_chess):
    g = golden_chess
    T, N, _ = (int(x) for x in g["shape"])
    C = torch.sparse_coo_tensor(torch.from_numpy(g["in_idx"].astype(np.int64)), torch.from_numpy(g["in_val"]),
                                (T, N, N)).coalesce()
    out = tg.func_MProduct(C, torch.from_numpy(g["M"]))
    assert torch.equal(out._indices().cpu(), torch.from_numpy(g["out_idx"].astype(np.int64)))
    np.testing.assert_allclose(out._values().cpu().numpy(), g["out_val"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("T,N,m,rho,b,norm", [(20, 3000, 9000, 0.9, 10, False), (9, 500, 4000, 0.0, 20, True),
                                               (6, 64, 2000, 0.5, 3, False), (33, 200, 300, 0.7, 32, False)])
def test_mtransform_sparse_vs_oracle(tg, T, N, m, rho, b, norm):
    from tmgcn_b200 import ops, synth
    idx, val = synth.synth_coo(N, T, m, rho, seed=T * 1000 + b)
    M = oracle.create_matrix_M(T, b, normalize=norm)
    ref_idx, ref_val = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy(), no_diag=b)
    band = tg.Band(M)
    # fp64 values (API parity) and fp32 values (pipeline layout)
    for dt, tol in ((torch.float64, 1e-13), (torch.float32, 2e-7)):
        csr = tg.SliceCSR.from_coo(idx, val, T, N, dtype=dt)
        out = ops.mtransform_sparse(csr, band)
        oi, ov = out.to_coo()
        assert torch.equal(oi.cpu(), torch.from_numpy(ref_idx))
        np.testing.assert_allclose(ov.cpu().double().numpy(), ref_val, rtol=tol, atol=0)
    # time-sharded: two ranks, halo = b-1 input slices from the predecessor
    cut = T // 2
    halo = min(b - 1, cut)
    sel_lo = idx[0] < cut
    sel_hi = idx[0] >= cut - halo
    lo = tg.SliceCSR.from_coo(idx[:, sel_lo], val[sel_lo], cut, N)
    idx_hi = idx[:, sel_hi].clone()
    idx_hi[0] -= cut - halo
    hi = tg.SliceCSR.from_coo(idx_hi, val[sel_hi], T - cut + halo, N)
    o_lo = ops.mtransform_sparse(lo, band, 0, cut, 0)
    o_hi = ops.mtransform_sparse(hi, band, cut, T, halo)
    i_lo, v_lo = o_lo.to_coo()
    i_hi, v_hi = o_hi.to_coo()
    i_hi = i_hi.clone()
    i_hi[0] += cut
    assert torch.equal(torch.cat([i_lo, i_hi], 1).cpu(), torch.from_numpy(ref_idx))
    np.testing.assert_allclose(torch.cat([v_lo, v_hi]).cpu().double().numpy(), ref_val, rtol=2e-7, atol=0)


def test_mtransform_sparse_fill_variants_agree(tg):
    """the fill-pass variants of the sparse M-transform (thread per row; shared-memory staged with 1, 2 or 4
    lanes per row; the automatic choice; the time-tiled kernels with the union-list fill and with the merging
    fill) in subprocesses, because the choice is latched on first use.  All of
    them sum a row's contributions in the same order, so indices AND values are bit-identical.  Hub rows
    overflow the staging capacity and take the kernel's global-memory path."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
import tmgcn_b200 as tg
from tmgcn_b200 import ops, synth
T, N = 14, 3000
torch.manual_seed(11)
idx, val = synth.synth_coo(N, T, 20000, 0.8, seed=3)
hub = torch.stack([torch.randint(0, T, (6000,)), torch.full((6000,), 7), torch.randint(0, N, (6000,))])
C = torch.sparse_coo_tensor(torch.cat([idx, hub], 1), torch.cat([val, torch.rand(6000, dtype=torch.float64)]), (T, N, N)).coalesce()
for b in (2, 5, 10, 24):
    band = tg.Band(tg.create_matrix_M(T, b))
    for dt in (torch.float32, torch.float64):
        out = ops.mtransform_sparse(tg.SliceCSR.from_coo(C._indices(), C._values(), T, N, dtype=dt), band)
        w = torch.arange(out.nnz, device=out.val.device, dtype=torch.float64) %% 97 + 1.0
        print(b, int(out.rowptr.sum()), int((out.col.to(torch.int64) * w.long()).sum()), out.nnz,
              repr(float((out.val.double() * w).sum())))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for flag in ("0", "1", "2", "4", None, "tiled", "untiled", "merging"):
        env = dict(os.environ)
        env.pop("TMGCN_MERGE_STAGED", None)
        env.pop("TMGCN_MERGE_TT", None)
        env.pop("TMGCN_MERGE_UNION", None)
        if flag == "tiled":
            env["TMGCN_MERGE_TT"] = "1"         # four output slices per merge pass (b <= 12): the default, with the
                                                # union-list fill (count pass records the pattern, fill never merges)
        elif flag == "merging":
            env["TMGCN_MERGE_UNION"] = "0"      # the same count pass, fill pass merges again (round-1/2 kernel)
        elif flag == "untiled":
            env["TMGCN_MERGE_TT"] = "0"         # the per-slice kernels, chosen by row length
        elif flag is not None:
            env["TMGCN_MERGE_STAGED"] = flag
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        outs[flag] = [ln.split() for ln in r.stdout.strip().splitlines()]
        assert len(outs[flag]) == 8
    for flag in ("1", "2", "4", None, "tiled", "untiled", "merging"):
        assert outs[flag] == outs["0"], flag


def test_mtransform_sparse_workspace_paths(tg):
    """the union-list variant through the raw C ABI with hand-sized workspaces: a record depth that most rows
    overflow (the run call must fall back to the merging fill), one that a few rows overflow (union fill + the
    overflow launch), the library's own size, and no workspace at all -- same bits every time.  fp64 values: the
    run call ignores the record (the union fill serves the fp32 layout only) and must still agree."""
    import ctypes as C
    from tmgcn_b200 import _lib, ops, synth
    from tmgcn_b200.ops import _p, _stream
    lib = _lib.load()
    T, N, b = 14, 3000, 5
    idx, val = synth.synth_coo(N, T, 3000, 0.8, seed=5)
    band = tg.Band(tg.create_matrix_M(T, b))
    w = band.device_weights(0, T, torch.float64)
    n_tasks = ((T + 3) // 4) * ((N + 31) // 32)
    for dt in (torch.float32, torch.float64):
        A = tg.SliceCSR.from_coo(idx, val, T, N, dtype=dt)
        ref = ops.mtransform_sparse(A, band)
        own = int(lib.tmgcn_mtransform_sparse_ws_bytes(T, 0, N, b, A.nnz))
        assert own > 256
        for ws_bytes in (0, 256 + n_tasks * (64 + 8 * 192), 256 + n_tasks * (64 + 13 * 192), own):
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=A.col.device) if ws_bytes else None
            counts = torch.empty(T * N, dtype=torch.int64, device=A.col.device)
            _lib.check(lib.tmgcn_mtransform_sparse_plan_ws(_p(A.rowptr), _p(A.col), T, 0, N, _p(w), b, _p(counts),
                                                           _p(ws), C.c_size_t(ws_bytes), _stream()))
            rowptr = ops.exclusive_scan(counts)
            assert torch.equal(rowptr, ref.rowptr), ws_bytes
            col = torch.full((ref.nnz,), -7, dtype=torch.int32, device=A.col.device)
            out = torch.full((ref.nnz,), float("nan"), dtype=dt, device=A.col.device)
            _lib.check(lib.tmgcn_mtransform_sparse_run_ws(_p(A.rowptr), _p(A.col), _p(A.val), T, 0, N, _p(w), b,
                                                          _p(rowptr), _p(col), _p(out), 1 if dt == torch.float64 else 0,
                                                          _p(ws), C.c_size_t(ws_bytes), _stream()))
            assert torch.equal(col, ref.col), ws_bytes
            assert torch.equal(out, ref.val), ws_bytes
            if ws is not None:
                print("workspace", str(dt), ws_bytes, "overflowed records:", int(ws[:4].view(torch.int32).item()))


def test_mtransform_sparse_zero_weight_and_cancellation(tg):
    """explicit zeros inside the band drop the source slice (nonzero(M[:, j]), ref: read_data.py:216);
    sums that cancel to 0.0 stay stored (coalesce never prunes)."""
    from tmgcn_b200 import ops
    T, N = 4, 5
    M = oracle.create_matrix_M(T, 3)
    M[2, 1] = 0.0
    M[3, 2] = -1.0
    idx = torch.tensor([[0, 1, 2, 2, 3], [0, 1, 1, 4, 1], [1, 2, 3, 4, 3]])
    val = torch.tensor([1.0, 2.0, 3.0, 4.0, 3.0], dtype=torch.float64)
    ref_idx, ref_val = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy())
    out = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N, dtype=torch.float64), tg.Band(M))
    oi, ov = out.to_coo()
    assert torch.equal(oi.cpu(), torch.from_numpy(ref_idx))
    np.testing.assert_allclose(ov.cpu().numpy(), ref_val, rtol=1e-14, atol=0)
    assert (ref_val == 0).any()      # the (3,1,3) entry: 1*3 + (-1)*3


def test_csr_transpose(tg):
    from tmgcn_b200 import synth
    T, N = 5, 400
    idx, val = synth.synth_coo(N, T, 3000, 0.5, seed=7)
    # make it unsymmetric and add a hub column (long transposed row)
    keep = torch.rand(idx.shape[1], generator=torch.Generator().manual_seed(1)) < 0.7
    idx, val = idx[:, keep], val[keep]
    hub = torch.stack([torch.full((N,), 2), torch.arange(N), torch.full((N,), 17)])
    idx = torch.cat([idx, hub], 1)
    val = torch.cat([val, torch.rand(N, dtype=torch.float64)])
    C = torch.sparse_coo_tensor(idx, val, (T, N, N)).coalesce()
    csr = tg.SliceCSR.from_coo(C._indices(), C._values(), T, N)
    tr = csr.transpose()
    ti, tv = tr.to_coo()
    Ct = C.transpose(1, 2).coalesce()
    assert torch.equal(ti.cpu(), Ct._indices())
    assert torch.equal(tv.cpu(), Ct._values().float())


# --------------------------------------------------------------------------
# (b) dense M-transform
# --------------------------------------------------------------------------
@pytest.mark.parametrize("T,b,NF", [(12, 1, 64), (12, 2, 96), (7, 3, 10), (30, 5, 1028), (25, 10, 4096),
                                    (40, 20, 260), (5, 20, 128), (70, 32, 36), (34, 20, 2001)])
@pytest.mark.parametrize("norm", [False, True])
def test_stencil_fwd_bwd(tg, T, b, NF, norm):
    from tmgcn_b200 import ops
    g = torch.Generator().manual_seed(T * 100 + b)
    M = oracle.create_matrix_M(T, b, normalize=norm)
    band = tg.Band(M)
    X = torch.rand(T, NF, 1, generator=g)
    G = torch.randn(T, NF, 1, generator=g)
    ref = (M @ X.double().reshape(T, -1)).reshape(X.shape)
    ref_g = (M.T @ G.double().reshape(T, -1)).reshape(X.shape)
    out = ops.stencil_fwd(X.cuda(), band)
    assert relerr(out, ref) <= 2e-6
    gin = ops.stencil_bwd(G.cuda(), band)
    assert relerr(gin, ref_g) <= 2e-6
    # sharded with halo: rank 1 owns [cut, T)
    cut = T // 2
    halo = min(band.b - 1, cut)
    out_hi = ops.stencil_fwd(X[cut - halo:].cuda().contiguous(), band, cut, T, halo)
    assert relerr(out_hi, ref[cut:]) <= 2e-6
    out_lo = ops.stencil_fwd(X[:cut].cuda().contiguous(), band, 0, cut, 0)
    assert relerr(out_lo, ref[:cut]) <= 2e-6
    g_hi = ops.stencil_bwd(G[cut:].cuda().contiguous(), band, cut, T, halo)      # (halo + T-cut) slices
    g_lo = ops.stencil_bwd(G[:cut].cuda().contiguous(), band, 0, cut, 0)
    total = torch.zeros(T, NF, 1, dtype=torch.float64)
    total[:cut] += g_lo.double().cpu()
    total[cut - halo:] += g_hi.double().cpu()        # halo part = what rank 1 owes rank 0
    assert relerr(total, ref_g) <= 2e-6
    # ranged variant: the halo slices first, the rest later, in place -- together identical to one call
    from tmgcn_b200 import _lib
    lib = _lib.load()
    Gh = G[cut:].cuda().contiguous()
    w = band.device_weights(cut, T, torch.float32)
    part = torch.full_like(g_hi, float("nan"))
    hb = min(band.b - 1, T - cut)
    _lib.check(lib.tmgcn_mtransform_dense_bwd_range(ops._p(Gh), ops._p(part), hb, halo, NF, ops._p(w), band.b, 0, halo,
                                                    -1, ops._stream()))
    assert torch.equal(part[:halo], g_hi[:halo]) and bool(torch.isnan(part[halo:]).all())
    _lib.check(lib.tmgcn_mtransform_dense_bwd_range(ops._p(Gh), ops._p(part), T - cut, halo, NF, ops._p(w), band.b, halo,
                                                    halo + T - cut, -1, ops._stream()))
    assert torch.equal(part, g_hi)
    # accumulate form: the last slices already hold what a successor rank sent; the stencil adds on top
    acc0 = halo + (T - cut) - min(2, T - cut)
    seeded = torch.zeros_like(g_hi)
    seeded[acc0:] = 0.25
    _lib.check(lib.tmgcn_mtransform_dense_bwd_range(ops._p(Gh), ops._p(seeded), T - cut, halo, NF, ops._p(w), band.b, 0,
                                                    halo + T - cut, acc0, ops._stream()))
    assert torch.equal(seeded[:acc0], g_hi[:acc0]) and torch.equal(seeded[acc0:], g_hi[acc0:] + 0.25)


# --------------------------------------------------------------------------
# (d) SpMM
# --------------------------------------------------------------------------
@pytest.mark.parametrize("F", [1, 2, 3, 6, 8, 20, 32, 100, 128, 256])
def test_spmm_vs_oracle(tg, F):
    from tmgcn_b200 import ops, synth
    T, N = 4, 700
    idx, val = synth.synth_coo(N, T, 5000, 0.6, seed=F)
    # hub rows: one with 40 entries, one dense row (> 32, multiple chunks)
    hub = torch.stack([torch.full((N,), 1), torch.full((N,), 5), torch.arange(N)])
    idx = torch.cat([idx, hub], 1)
    val = torch.cat([val, torch.rand(N, dtype=torch.float64)])
    C = torch.sparse_coo_tensor(idx, val, (T, N, N)).coalesce()
    A = oracle.split_slices(C._indices().numpy(), C._values().numpy(), T, N)
    X = torch.rand(T, N, F, generator=torch.Generator().manual_seed(F), dtype=torch.float64)
    ref = oracle.compute_AX(A, X)
    csr = tg.SliceCSR.from_coo(C._indices(), C._values(), T, N)
    out = ops.spmm_raw(csr, X.float().cuda())
    assert relerr(out, ref) <= TOL_OUT
    # transposed (backward) product
    ref_t = oracle.compute_AX([a.t().coalesce() for a in A], X)
    out_t = ops.spmm_raw(csr.transpose(), X.float().cuda())
    assert relerr(out_t, ref_t) <= TOL_OUT


@pytest.mark.parametrize("F,act", [(128, "none"), (64, "relu"), (96, "selu"), (256, "none")])
def test_spmm_short_rows(tg, F, act):
    """graphs with few stored entries per row (the C1-C4 shapes): ragged rows (many empty ones, one 70-entry hub)
    against the oracle, forward and transposed, with the activation epilogues."""
    from tmgcn_b200 import ops
    T, N = 5, 900
    g = torch.Generator().manual_seed(F)
    nnz = 2600
    idx = torch.stack([torch.randint(0, T, (nnz,), generator=g), torch.randint(0, N, (nnz,), generator=g),
                       torch.randint(0, N, (nnz,), generator=g)])
    idx[1, :300] = torch.randint(0, 40, (300,), generator=g)       # leaves many rows empty
    hub = torch.stack([torch.full((70,), 2), torch.full((70,), 11), torch.arange(0, 140, 2)])
    idx = torch.cat([idx, hub], 1)
    C = torch.sparse_coo_tensor(idx, torch.rand(idx.shape[1], dtype=torch.float64, generator=g) - 0.3, (T, N, N)).coalesce()
    assert C._nnz() < 8 * T * N
    A = oracle.split_slices(C._indices().numpy(), C._values().numpy(), T, N)
    X = torch.rand(T, N, F, generator=g, dtype=torch.float64) - 0.5
    ref = oracle.nonlin(act)(oracle.compute_AX(A, X).double())
    csr = tg.SliceCSR.from_coo(C._indices(), C._values(), T, N)
    out = ops.spmm_raw(csr, X.float().cuda(), tg.ops.ACT[act])
    assert relerr(out, ref) <= TOL_OUT
    ref_t = oracle.compute_AX([a.t().coalesce() for a in A], X)
    assert relerr(ops.spmm_raw(csr.transpose(), X.float().cuda()), ref_t) <= TOL_OUT


@pytest.mark.parametrize("act", ["relu", "leaky", "selu"])
def test_spmm_activation_and_grad(tg, act):
    from tmgcn_b200 import ops, synth
    T, N, F = 3, 300, 12
    idx, val = synth.synth_coo(N, T, 2000, 0.5, seed=3)
    A = oracle.split_slices(idx.numpy(), val.numpy(), T, N)
    csr = tg.SliceCSR.from_coo(idx, val, T, N)
    g = torch.Generator().manual_seed(5)
    X = (torch.rand(T, N, F, generator=g, dtype=torch.float64) - 0.5)
    G = torch.randn(T, N, F, generator=g)
    Xr = X.clone().requires_grad_(True)
    ref = oracle.nonlin(act)(torch.stack([torch.sparse.mm(A[k], Xr[k]) for k in range(T)]))
    ref.backward(G.double())
    Xd = X.float().cuda().requires_grad_(True)
    out = ops.spmm(csr, Xd, act)
    out.backward(G.cuda())
    assert relerr(out, ref) <= TOL_OUT
    assert relerr(Xd.grad, Xr.grad) <= TOL_GRAD


# --------------------------------------------------------------------------
# (c) GEMM
# --------------------------------------------------------------------------
@pytest.mark.parametrize("R,K,Nf", [(1000, 2, 6), (777, 6, 6), (500, 6, 2), (300, 32, 48), (2000, 128, 128),
                                    (130, 128, 128), (4100, 128, 64), (257, 100, 36), (640, 256, 128)])
@pytest.mark.parametrize("act", ["none", "relu", "leaky", "selu"])
def test_gemm_fwd_bwd(tg, R, K, Nf, act):
    from tmgcn_b200 import ops
    g = torch.Generator().manual_seed(R + K)
    P = torch.randn(R, K, generator=g)
    W = torch.randn(K, Nf, generator=g) / K ** 0.5
    G = torch.randn(R, Nf, generator=g)
    Pr, Wr = P.double().requires_grad_(True), W.double().requires_grad_(True)
    ref = oracle.nonlin(act)(Pr @ Wr)
    ref.backward(G.double())
    Pd, Wd = P.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
    out = ops.gemm_xw(Pd, Wd, act)
    out.backward(G.cuda())
    assert relerr(out, ref) <= TOL_OUT
    assert relerr(Pd.grad, Pr.grad) <= TOL_GRAD
    assert relerr(Wd.grad, Wr.grad) <= TOL_GRAD


# --------------------------------------------------------------------------
# (e) edge readout
# --------------------------------------------------------------------------
@pytest.mark.parametrize("F,Cc", [(2, 2), (6, 3), (16, 1), (128, 2), (100, 8), (12, 11), (64, 40)])
def test_edge_readout_and_gather(tg, F, Cc):
    from tmgcn_b200 import ops
    T, N, E = 5, 60, 700
    g = torch.Generator().manual_seed(F)
    Y = torch.randn(T, N, F, generator=g)
    U = torch.randn(2 * F, Cc, generator=g)
    edges = torch.stack([torch.randint(0, T, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                         torch.randint(0, N, (E,), generator=g)])
    edges[:, :40] = edges[:, :1]          # heavy duplicates: the scatter-add must accumulate them
    dOut = torch.randn(E, Cc, generator=g)
    dZ = torch.randn(E, 2 * F, generator=g)
    src, trg = oracle.flat_edge_ids(edges, N)
    Yr, Ur = Y.double().requires_grad_(True), U.double().requires_grad_(True)
    Zr = torch.cat((Yr.reshape(-1, F)[src], Yr.reshape(-1, F)[trg]), 1)
    ref = Zr @ Ur
    ref.backward(dOut.double())
    plan = tg.EdgePlan(edges, N)
    assert torch.equal(plan.src.cpu(), src) and torch.equal(plan.dst.cpu(), trg)
    Yd, Ud = Y.cuda().requires_grad_(True), U.cuda().requires_grad_(True)
    out = ops.edge_readout(Yd, Ud, plan)
    out.backward(dOut.cuda())
    assert relerr(out, ref) <= TOL_OUT
    assert relerr(Yd.grad, Yr.grad) <= TOL_GRAD
    assert relerr(Ud.grad, Ur.grad) <= TOL_GRAD
    # plain gather (E, 2F) and its scatter-add backward
    Yr2 = Y.double().requires_grad_(True)
    Zr2 = torch.cat((Yr2.reshape(-1, F)[src], Yr2.reshape(-1, F)[trg]), 1)
    Zr2.backward(dZ.double())
    Yd2 = Y.cuda().requires_grad_(True)
    Z = ops.edge_gather(Yd2, plan)
    Z.backward(dZ.cuda())
    assert torch.equal(Z.cpu(), Zr2.detach().float())       # pure data movement: exact
    assert relerr(Yd2.grad, Yr2.grad) <= TOL_GRAD
    # determinism of the scatter-add
    Yd3 = Y.cuda().requires_grad_(True)
    ops.edge_gather(Yd3, plan).backward(dZ.cuda())
    assert torch.equal(Yd3.grad, Yd2.grad)


@pytest.mark.parametrize("act", ["none", "relu", "leaky", "selu"])
def test_activation(tg, act):
    from tmgcn_b200 import ops
    x = torch.randn(10_001, generator=torch.Generator().manual_seed(1))
    xr = x.double().requires_grad_(True)
    ref = oracle.nonlin(act)(xr)
    ref.backward(torch.ones_like(ref))
    xd = x.cuda().requires_grad_(True)
    out = ops.activation(xd, act)
    out.backward(torch.ones_like(out))
    assert relerr(out, ref) <= 1e-6 and relerr(xd.grad, xr.grad) <= 1e-6


# --------------------------------------------------------------------------
# modules against the reference's own outputs (golden) and the oracle
# --------------------------------------------------------------------------
def _inputs(g):
    T, N = (int(x) for x in g["TN"])
    M = torch.from_numpy(g["M"])
    At = oracle.split_slices(g["Ct_idx"], g["Ct_val"], T, N)
    A = oracle.split_slices(g["C_idx"], g["C_val"], T, N)
    X, X2 = torch.from_numpy(g["X"]), torch.from_numpy(g["X2"])
    edges, edges2 = torch.from_numpy(g["edges"]), torch.from_numpy(g["edges2"])
    return T, N, M, At, A, X, X2, edges, edges2


def _load(mod, g, prefix, names):
    with torch.no_grad():
        for n in names:
            getattr(mod, n).copy_(torch.from_numpy(g[prefix + n]))


def test_module_gcn1_golden(tg, golden_models):
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    m = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    _load(m, g, "gcn1_", ["W", "U"])
    assert relerr(m.AtXt, g["gcn1_AtXt"]) <= TOL_OUT
    out = m()
    assert relerr(out, g["gcn1_out"]) <= TOL_OUT
    assert_same_argmax(out, g["gcn1_out"])
    out.backward(torch.from_numpy(g["gcn1_dOut"]).cuda())
    assert relerr(m.W.grad, g["gcn1_dW"]) <= TOL_GRAD
    assert relerr(m.U.grad, g["gcn1_dU"]) <= TOL_GRAD
    with torch.no_grad():
        fresh = m(At, X2, edges2)
        assert relerr(fresh, g["gcn1_out_fresh"]) <= TOL_OUT
        # cached == recomputed, bit for bit (SURVEY section 4 invariant 2)
        assert torch.equal(m(At, X, edges), m())


@pytest.mark.parametrize("tag,kw", [
    ("relu", dict(nonlin2="relu")), ("leaky", dict(nonlin2="leaky")), ("selu", dict(nonlin2="selu")),
    ("selu_m2", dict(nonlin2="selu", apply_M_twice=True)),
    ("relu_m3", dict(nonlin2="relu", apply_M_twice=True, apply_M_three_times=True))])
def test_module_gcn2_golden(tg, golden_models, tag, kw):
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    p = "gcn2_" + tag + "_"
    m = tg.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=True, use_Minv=False, **kw)
    _load(m, g, p, ["W1", "W2", "U"])
    out = m()
    assert relerr(out, g[p + "out"]) <= TOL_OUT
    assert_same_argmax(out, g[p + "out"])
    out.backward(torch.from_numpy(g["gcn1_dOut"]).cuda())
    for n in ("W1", "W2", "U"):
        assert relerr(getattr(m, n).grad, g[p + "d" + n]) <= TOL_GRAD, n
    with torch.no_grad():
        assert relerr(m(At, X2, edges2), g[p + "out_fresh"]) <= TOL_OUT


@pytest.mark.parametrize("tag,hf", [("kw1", [6, 2]), ("kw2", [6, 6, 2])])
def test_module_kwgcn_golden(tg, golden_models, tag, hf):
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    p = tag + "_"
    m = tg.EmbeddingKWGCN(A, X, edges, hidden_feat=hf, nonlin2="leaky")
    names = ["W1", "U"] + (["W2"] if len(hf) == 3 else [])
    _load(m, g, p, names)
    out = m()
    assert relerr(out, g[p + "out"]) <= TOL_OUT
    out.backward(torch.from_numpy(g["gcn1_dOut"]).cuda())
    for n in names:
        assert relerr(getattr(m, n).grad, g[p + "d" + n]) <= TOL_GRAD, n
    with torch.no_grad():
        assert relerr(m(A, X2, edges2), g[p + "out_fresh"]) <= TOL_OUT


def test_module_wide_golden(tg, golden_models):
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    m = tg.EmbeddingGCN2(At, torch.from_numpy(g["wide_X"]), edges, M, hidden_feat=[48, 16, 3], condensed_W=True,
                         use_Minv=False, apply_M_twice=True, nonlin2="relu")
    _load(m, g, "wide_", ["W1", "W2", "U"])
    out = m()
    assert relerr(out, g["wide_out"]) <= TOL_OUT
    out.backward(torch.from_numpy(g["wide_dOut"]).cuda())
    for n in ("W1", "W2", "U"):
        assert relerr(getattr(m, n).grad, g["wide_d" + n]) <= TOL_GRAD, n


def test_module_seed_reproduces_reference_init(tg, golden_models):
    """same torch seed => same randn draws in the reference's order (ehf:188-192)."""
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    torch.manual_seed(100)
    m = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    assert torch.equal(m.W.detach().cpu(), torch.from_numpy(g["gcn1_W"]))
    assert torch.equal(m.U.detach().cpu(), torch.from_numpy(g["gcn1_U"]))
    torch.manual_seed(300)
    k = tg.EmbeddingKWGCN(A, X, edges, hidden_feat=[6, 6, 2], nonlin2="leaky")
    assert torch.equal(k.W2.detach().cpu(), torch.from_numpy(g["kw2_W2"]))
    assert torch.equal(k.W1.detach().cpu(), torch.from_numpy(g["kw2_W1"]))


def test_module_per_slice_weights_and_regression_head(tg, golden_models):
    """condensed_W=False (ehf:188-191, 277-282) and EmbeddingGCN_reg (ehf:359-423) against the reference."""
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    dOut = torch.from_numpy(g["gcn1_dOut"]).cuda()
    m = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=False, use_Minv=False)
    assert tuple(m.W.shape) == (T, 2, 6)
    _load(m, g, "gcn1u_", ["W", "U"])
    out = m()
    assert relerr(out, g["gcn1u_out"]) <= TOL_OUT
    out.backward(dOut)
    assert relerr(m.W.grad, g["gcn1u_dW"]) <= TOL_GRAD and relerr(m.U.grad, g["gcn1u_dU"]) <= TOL_GRAD
    m2 = tg.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=False, use_Minv=False,
                          apply_M_twice=True, nonlin2="leaky")
    _load(m2, g, "gcn2u_", ["W1", "W2", "U"])
    out = m2()
    assert relerr(out, g["gcn2u_out"]) <= TOL_OUT
    out.backward(dOut)
    for n in ("W1", "W2", "U"):
        assert relerr(getattr(m2, n).grad, g["gcn2u_d" + n]) <= TOL_GRAD, n
    r = tg.EmbeddingGCN_reg(At, X, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    with torch.no_grad():
        r.W.copy_(torch.from_numpy(g["reg_W"]))
        r.lin1.weight.copy_(torch.from_numpy(g["reg_lw"]))
        r.lin1.bias.copy_(torch.from_numpy(g["reg_lb"]))
    out = r()
    assert tuple(out.shape) == (T, N) and relerr(out, g["reg_out"]) <= TOL_OUT
    out.backward(torch.from_numpy(g["reg_dOut"]).cuda())
    assert relerr(r.W.grad, g["reg_dW"]) <= TOL_GRAD
    assert relerr(r.lin1.weight.grad, g["reg_dlw"]) <= TOL_GRAD and relerr(r.lin1.bias.grad, g["reg_dlb"]) <= TOL_GRAD


@pytest.mark.parametrize("K,Nf,act", [(2, 6, "none"), (6, 6, "selu"), (70, 33, "relu"), (128, 128, "leaky")])
def test_gemm_xw_sliced_grouped_launch(tg, K, Nf, act):
    """condensed_W=False (ehf:188-191): y[t] = act(p[t] @ w[t]) for all slices in one grouped launch, with
    dP and per-slice dW, against torch fp64."""
    from tmgcn_b200 import _lib, ops
    T, N = 7, 150
    g = torch.Generator().manual_seed(K + Nf)
    p = torch.randn(T, N, K, generator=g)
    w = torch.randn(T, K, Nf, generator=g) / K ** 0.5
    dy = torch.randn(T, N, Nf, generator=g)
    pr, wr = p.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = oracle.nonlin(act)(torch.matmul(pr, wr))
    ref.backward(dy.double())
    pd, wd = p.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    n0 = _lib.launch_count()
    out = ops.gemm_xw_sliced(pd, wd, act)
    n_fwd = _lib.launch_count() - n0
    out.backward(dy.cuda())
    if not (K == 128 and Nf == 128):
        assert n_fwd == 1                         # one grouped launch, not one per slice
    assert relerr(out, ref) <= TOL_OUT
    assert relerr(pd.grad, pr.grad) <= TOL_GRAD and relerr(wd.grad, wr.grad) <= TOL_GRAD


def test_module_use_minv(tg, golden_models):
    """use_Minv=True (ehf:183-184, 223-224, 331-341): inv(M) applied as a banded substitution."""
    from tmgcn_b200 import ops
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    dOut = torch.from_numpy(g["gcn1_dOut"]).cuda()
    # the substitution against the dense inverse, forward and adjoint, on a longer band
    T2, b2 = 60, 20
    M2 = oracle.create_matrix_M(T2, b2)
    band2 = tg.Band(M2)
    Z = torch.randn(T2, 37, 5, generator=torch.Generator().manual_seed(1))
    Minv = torch.linalg.inv(M2)
    Zd = Z.cuda().requires_grad_(True)
    Yd = ops.mtransform_dense_inv(Zd, band2)
    assert relerr(Yd, (Minv @ Z.double().reshape(T2, -1)).reshape(Z.shape)) <= TOL_OUT
    G = torch.randn(Z.shape, generator=torch.Generator().manual_seed(2))
    Yd.backward(G.cuda())
    assert relerr(Zd.grad, (Minv.T @ G.double().reshape(T2, -1)).reshape(Z.shape)) <= TOL_GRAD
    # modules: the reference raises a dtype error on this flag (see oracle.apply_Minv), so the check is
    # against the oracle's dense-inverse restatement with the same weights
    m = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=True)
    ref = oracle.OracleGCN(At, X, edges, M, m.W.detach().cpu(), m.U.detach().cpu(), as_reference=False,
                           use_Minv=True)
    out, out_r = m(), ref()
    assert relerr(out, out_r) <= TOL_OUT
    out.backward(dOut)
    out_r.backward(dOut.cpu())
    assert relerr(m.W.grad, ref.W.grad) <= TOL_GRAD and relerr(m.U.grad, ref.U.grad) <= TOL_GRAD
    m2 = tg.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=True, use_Minv=True, nonlin2="selu")
    ref2 = oracle.OracleGCN2(At, X, edges, M, m2.W1.detach().cpu(), m2.W2.detach().cpu(), m2.U.detach().cpu(),
                             nonlin2="selu", as_reference=False, use_Minv=True)
    out, out_r = m2(), ref2()
    assert relerr(out, out_r) <= TOL_OUT
    out.backward(dOut)
    out_r.backward(dOut.cpu())
    for n in ("W1", "W2", "U"):
        assert relerr(getattr(m2, n).grad, getattr(ref2, n).grad) <= TOL_GRAD, n


@pytest.mark.parametrize("transposed", [False, True])
def test_solve_part_chained_blocks(tg, transposed):
    """the time-sharded / column-chunked substitution kernel (tmgcn_mtransform_dense_solve_part): three uneven
    time blocks chained on one GPU through their halo rows, three column chunks each, against inv(M)."""
    from tmgcn_b200 import sharding
    T, b, N, F = 41, 6, 33, 5
    M = oracle.create_matrix_M(T, b, normalize=True)
    band = tg.Band(M)
    Z = torch.randn(T, N, F, generator=torch.Generator().manual_seed(4))
    Minv = torch.linalg.inv(M)
    want = ((Minv.T if transposed else Minv) @ Z.double().reshape(T, -1)).reshape(Z.shape)
    blocks = [(0, 9), (9, 27), (27, 41)]
    order = blocks[::-1] if transposed else blocks
    NF, h = N * F, b - 1
    out = torch.empty(T, NF, device="cuda")
    Zd = Z.cuda().reshape(T, NF)
    halo = None
    for t0, t1 in order:
        Tl = t1 - t0
        rows_w = min(T, t1 + h) if transposed else t1
        w = band.device_weights(t0, rows_w, torch.float32)
        z, y = Zd[t0:t1].contiguous(), torch.empty(Tl, NF, device="cuda")
        for c in range(3):
            c0, c1 = NF * c // 3, NF * (c + 1) // 3
            hk = 0 if halo is None else h
            hc = None if halo is None else halo[:, c0:c1].contiguous()
            sharding._device_local_solve(z[:, c0:], y[:, c0:], hc, Tl, hk, c1 - c0, NF, c1 - c0, w, b, transposed)
        out[t0:t1] = y
        halo = y[:h].clone() if transposed else y[Tl - h:].clone()
    assert relerr(out.reshape(T, N, F), want) <= TOL_OUT


def test_module_use_minv_golden(tg, golden_minv):
    """use_Minv=True against the UNMODIFIED reference (run on all-fp32 inputs, the one dtype configuration in
    which it executes the flag: tests/golden/make_golden.py::gen_minv): 1-layer with shared and per-slice
    weights (ehf:222-224) and the regression head (ehf:415-417)."""
    g = golden_minv
    T, N = (int(x) for x in g["TN"])
    M, X, edges = torch.from_numpy(g["M"]), torch.from_numpy(g["X"]), torch.from_numpy(g["edges"])
    At = [a.coalesce() for a in oracle.split_slices(g["Ct_idx"], g["Ct_val"], T, N)]
    dOut = torch.from_numpy(g["dOut"]).cuda()
    for tag, cw in (("gcn1", True), ("gcn1u", False)):
        m = tg.EmbeddingGCN(At, X, edges, M, hidden_feat=[5, 2], condensed_W=cw, use_Minv=True)
        _load(m, g, tag + "_", ["W", "U"])
        out = m()
        assert relerr(out, g[tag + "_out"]) <= TOL_OUT, tag
        out.backward(dOut)
        assert relerr(m.W.grad, g[tag + "_dW"]) <= TOL_GRAD and relerr(m.U.grad, g[tag + "_dU"]) <= TOL_GRAD, tag
    r = tg.EmbeddingGCN_reg(At, X, M, hidden_feat=[5, 2], condensed_W=True, use_Minv=True)
    with torch.no_grad():
        r.W.copy_(torch.from_numpy(g["reg_W"]))
        r.lin1.weight.copy_(torch.from_numpy(g["reg_lw"]))
        r.lin1.bias.copy_(torch.from_numpy(g["reg_lb"]))
    out = r()
    assert relerr(out, g["reg_out"]) <= TOL_OUT
    out.backward(torch.from_numpy(g["reg_dOut"]).cuda())
    assert relerr(r.W.grad, g["reg_dW"]) <= TOL_GRAD


def test_module_errors(tg, golden_models):
    g = golden_models
    T, N, M, At, A, X, X2, edges, edges2 = _inputs(g)
    with pytest.raises(AssertionError):
        tg.func_MProduct(torch.eye(3).reshape(1, 3, 3).to_sparse(), torch.eye(2, dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        tg.Band(torch.ones(4, 4))
    with pytest.raises(RuntimeError):
        from tmgcn_b200 import ops
        ops.stencil_fwd(torch.zeros(7, 8, 1).cuda(), tg.Band(oracle.create_matrix_M(4, 3)), 0, 4, 3)  # halo > b-1


# --------------------------------------------------------------------------
# the benchmarked layer at a mid size, F = 128 (tensor-core GEMM path)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["none", "relu"])
def test_layer_f128_vs_oracle(tg, act):
    from tmgcn_b200 import ops, synth
    T, N, F, Cc, b = 12, 2000, 128, 2, 10
    idx, val = synth.synth_coo(N, T, 8000, 0.9, seed=99)
    M = oracle.create_matrix_M(T, b)
    band = tg.Band(M)
    ref_idx, ref_val = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy())
    At_ref = oracle.split_slices(ref_idx, ref_val, T, N)
    At = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx, val, T, N), band)
    g = torch.Generator().manual_seed(1)
    H = torch.rand(T, N, F, generator=g)
    W = torch.randn(F, F, generator=g) / F ** 0.5
    U = torch.randn(2 * F, Cc, generator=g)
    E = 5000
    edges = synth.synth_edges(At, E).cpu()
    dOut = torch.randn(E, Cc, generator=g)
    out_r, dH_r, dW_r, dU_r = oracle.layer_fwd_bwd(At_ref, H, M, W, U, edges, dOut, act, as_reference=False)
    layer = tg.TMGCNLayer(At, band, tg.EdgePlan(edges, N), W, U, act)
    Hd = H.cuda().requires_grad_(True)
    out = layer(Hd)
    out.backward(dOut.cuda())
    assert relerr(out, out_r) <= TOL_OUT
    assert_same_argmax(out, out_r)
    assert relerr(Hd.grad, dH_r) <= TOL_GRAD
    assert relerr(layer.W.grad, dW_r) <= TOL_GRAD
    assert relerr(layer.U.grad, dU_r) <= TOL_GRAD


# --------------------------------------------------------------------------
# the workspace-planned step bench.py times == the autograd path, also when sharded
# --------------------------------------------------------------------------
@pytest.mark.parametrize("act,mode", [("none", "dense"), ("none", "lowrank"), ("selu", "auto")])
def test_layer_step_matches_autograd_and_sharding(tg, act, mode):
    from tmgcn_b200 import ops, synth
    from tmgcn_b200.layer_step import LayerStep
    T, N, F, Cc, b = 10, 900, 32, 3, 4
    idx, val = synth.synth_coo(N, T, 4000, 0.8, seed=5)
    M = oracle.create_matrix_M(T, b)
    band = tg.Band(M)
    A = tg.SliceCSR.from_coo(idx, val, T, N)
    At = ops.mtransform_sparse(A, band)
    g = torch.Generator().manual_seed(2)
    H = torch.rand(T, N, F, generator=g).cuda()
    W = (torch.randn(F, F, generator=g) / F ** 0.5).cuda()
    U = torch.randn(2 * F, Cc, generator=g).cuda()
    E = 3000
    edges = synth.synth_edges(At, E)
    dOut = torch.randn(E, Cc, generator=g).cuda()
    plan = tg.EdgePlan(edges, N)
    layer = tg.TMGCNLayer(At, band, plan, W, U, act)
    Hd = H.clone().requires_grad_(True)
    out_ref = layer(Hd)
    out_ref.backward(dOut)
    step = LayerStep(At, band, plan, F, F, Cc, act, bwd_mode=mode)
    assert step.bwd_mode == ("dense" if act != "none" else mode)
    out = step.forward(H, W, U)
    assert torch.equal(out, out_ref.detach())
    dH, dW, dU = step.backward(dOut, W, U)
    if step.bwd_mode == "dense":      # same kernels as the autograd path: bit-identical
        assert torch.equal(dH, Hd.grad) and torch.equal(dW, layer.W.grad) and torch.equal(dU, layer.U.grad)
    else:                             # low-rank association: same gradients to rounding
        assert relerr(dH, Hd.grad) <= TOL_GRAD and relerr(dW, layer.W.grad) <= TOL_GRAD
        assert relerr(dU, layer.U.grad) <= TOL_GRAD
        with pytest.raises(ValueError):
            LayerStep(At, band, plan, F, F, Cc, "relu", bwd_mode="lowrank")
    # two time shards: rank 1 owns [cut, T) with a (b-1)-slice halo
    cut, halo = 6, b - 1
    sel = idx[0] >= cut - halo
    idx_hi = idx[:, sel].clone()
    idx_hi[0] -= cut - halo
    At_hi = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx_hi, val[sel], T - cut + halo, N), band, cut, T, halo)
    e_hi = edges[:, edges[0] >= cut]
    plan_hi = tg.EdgePlan(e_hi, N, t_offset=cut)
    step_hi = LayerStep(At_hi, band, plan_hi, F, F, Cc, act, cut, T, halo, bwd_mode=mode)
    out_hi = step_hi.forward(H[cut - halo:].contiguous(), W, U)
    assert relerr(out_hi, out_ref.detach()[edges[0].cpu() >= cut]) <= TOL_OUT
    sel_e = (edges[0] >= cut)
    dH_hi, dW_hi, dU_hi = step_hi.backward(dOut[sel_e].contiguous(), W, U)
    e_lo = edges[:, ~sel_e]
    sel_lo = idx[0] < cut
    At_lo = ops.mtransform_sparse(tg.SliceCSR.from_coo(idx[:, sel_lo], val[sel_lo], cut, N), band, 0, cut, 0)
    step_lo = LayerStep(At_lo, band, tg.EdgePlan(e_lo, N), F, F, Cc, act, 0, cut, 0, bwd_mode=mode)
    step_lo.forward(H[:cut].contiguous(), W, U)
    dH_lo, dW_lo, dU_lo = step_lo.backward(dOut[~sel_e].contiguous(), W, U)
    total = torch.zeros_like(H)
    total[:cut] += dH_lo
    total[cut - halo:] += dH_hi
    assert relerr(total, Hd.grad) <= TOL_GRAD
    assert relerr(dW_lo + dW_hi, layer.W.grad) <= TOL_GRAD
    assert relerr(dU_lo + dU_hi, layer.U.grad) <= TOL_GRAD


# --------------------------------------------------------------------------
# graph preparation (ref: read_data.py:88-188)
# --------------------------------------------------------------------------
def test_graph_preparation_golden_and_oracle(tg):
    import os
    from conftest import GOLDEN
    from tmgcn_b200 import preprocess as pp
    g = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    TT, N, w = (int(x) for x in g["TT_N_w"])

    def coo(name):
        return torch.sparse_coo_tensor(torch.from_numpy(g[name + "_idx"]), torch.from_numpy(g[name + "_val"]),
                                       (TT, N, N)).coalesce()

    def same(x, name, rtol):
        assert x.is_coalesced() and x.dtype == torch.float64
        assert torch.equal(x._indices().cpu(), torch.from_numpy(g[name + "_idx"]))
        np.testing.assert_allclose(x._values().cpu().numpy(), g[name + "_val"], rtol=rtol, atol=0)
    B = pp.func_make_symmetric(coo("A"), N, TT)
    same(B, "sym", 1e-15)
    E = pp.func_edge_life(B, N, TT, edge_life_window=w)
    same(E, "life", 1e-15)
    C = pp.func_laplacian_transformation(E, N, TT)
    same(C, "lap", 1e-14)
    S = pp.func_create_sparse(C, N, TT, 4, 2, 6)
    assert tuple(S.shape) == (4, N, N)
    assert torch.equal(S._indices().cpu(), torch.from_numpy(g["win_idx"]))
    np.testing.assert_allclose(S._values().cpu().numpy(), g["win_val"], rtol=1e-14, atol=0)
    # a bigger unsymmetric random tensor with self loops, duplicates and empty slices against the oracle
    T2, N2 = 9, 2500
    gen = torch.Generator().manual_seed(4)
    n = 60000
    idx = torch.stack([torch.randint(1, T2 - 1, (n,), generator=gen), torch.randint(0, N2, (n,), generator=gen),
                       torch.randint(0, N2, (n,), generator=gen)])
    A2 = torch.sparse_coo_tensor(idx, torch.ones(n, dtype=torch.float64), (T2, N2, N2)).coalesce()
    ai, av = A2._indices().numpy(), A2._values().numpy()
    si, sv = oracle.make_symmetric(ai, av, T2, N2)
    li, lv = oracle.edge_life(si, sv, T2, N2, 4)
    ci, cv = oracle.laplacian_transformation(li, lv, T2, N2)
    out = pp.func_laplacian_transformation(pp.func_edge_life(pp.func_make_symmetric(A2, N2, T2), N2, T2, 4), N2, T2)
    assert torch.equal(out._indices().cpu(), torch.from_numpy(ci))
    np.testing.assert_allclose(out._values().cpu().numpy(), cv, rtol=1e-13, atol=0)
    # the synthetic-input generator's preparation (symmetrise + I + normalise) is the same pipeline
    ni, nv = oracle.normalise_adjacency(ai, av, T2, N2)
    out2 = pp.func_laplacian_transformation(pp.func_make_symmetric(A2, N2, T2), N2, T2)
    assert torch.equal(out2._indices().cpu(), torch.from_numpy(ni))
    np.testing.assert_allclose(out2._values().cpu().numpy(), nv, rtol=1e-13, atol=0)


# --------------------------------------------------------------------------
# degenerate shapes through every stage of the path
# --------------------------------------------------------------------------
def test_degenerate_inputs(tg):
    """empty tensors, a single node / slice, band 1, no edges, odd feature widths: every stage returns the
    right shapes and the oracle's numbers instead of tripping over a zero-sized launch"""
    from tmgcn_b200 import ops
    dev = "cuda"
    # (a) empty sparse tensor, and T = N = 1
    M3 = tg.create_matrix_M(3, 2)
    empty = torch.sparse_coo_tensor(torch.zeros(3, 0, dtype=torch.int64), torch.zeros(0, dtype=torch.float64), (3, 4, 4))
    out = tg.func_MProduct(empty.coalesce(), M3, no_diag=2)
    assert out._nnz() == 0 and tuple(out.shape) == (3, 4, 4)
    one = torch.sparse_coo_tensor(torch.zeros(3, 1, dtype=torch.int64), torch.tensor([2.5], dtype=torch.float64), (1, 1, 1))
    out = tg.func_MProduct(one.coalesce(), tg.create_matrix_M(1, 1), no_diag=1)
    assert out._nnz() == 1 and float(out._values()[0]) == 2.5
    # (b) stencil with b = 1 is the identity scaled by the diagonal; T = 1
    x = torch.randn(1, 5, 3, device=dev)
    y = ops.stencil_fwd(x, tg.Band(tg.create_matrix_M(1, 1)))
    assert torch.equal(y, x)
    # (d) SpMM over an all-empty CSR and odd widths (F = 1, 3, 6: the scalar / float2 paths)
    T, N = 2, 6
    csr0 = tg.SliceCSR.from_coo(torch.zeros(3, 0, dtype=torch.int64), torch.zeros(0), T, N)
    for F in (1, 3, 6):
        X = torch.randn(T, N, F, device=dev)
        assert torch.count_nonzero(ops.spmm_raw(csr0, X)) == 0
        idx = torch.tensor([[0, 0, 1, 1, 1], [0, 5, 2, 2, 3], [1, 5, 0, 4, 3]])
        val = torch.tensor([1.0, -2.0, 0.5, 4.0, 3.0])
        csr = tg.SliceCSR.from_coo(idx, val, T, N)
        ref = torch.zeros(T, N, F, dtype=torch.float64)
        for (t, i, j), v in zip(idx.t().tolist(), val.tolist()):
            ref[t, i] += v * X[t, j].double().cpu()
        assert relerr(ops.spmm_raw(csr, X), ref) < 1e-6
    # (c) GEMM with zero rows and a 1 x 1 weight
    W = torch.randn(3, 2, device=dev)
    assert tuple(ops.gemm_xw(torch.zeros(0, 4, 3, device=dev), W).shape) == (0, 4, 2)
    assert relerr(ops.gemm_xw(torch.full((1, 1, 1), 3.0, device=dev), torch.full((1, 1), -2.0, device=dev)),
                  torch.full((1, 1, 1), -6.0)) < 1e-7
    # (e) readout with no edges: empty logits, zero gradients of the right shapes
    Y = torch.randn(2, 6, 4, device=dev, requires_grad=True)
    U = torch.randn(8, 2, device=dev, requires_grad=True)
    plan = tg.EdgePlan(torch.zeros(3, 0, dtype=torch.int64), 6)
    logits = ops.edge_readout(Y, U, plan)
    assert tuple(logits.shape) == (0, 2)
    logits.sum().backward()
    assert torch.count_nonzero(Y.grad) == 0 and torch.count_nonzero(U.grad) == 0
    # a model on a graph whose first and last slices are empty
    idx = torch.tensor([[1, 1, 2], [0, 3, 2], [3, 0, 2]])
    At = tg.SliceCSR.from_coo(idx, torch.tensor([0.5, 0.5, 1.0]), 4, 5)
    X = torch.randn(4, 5, 2)
    edges = torch.tensor([[0, 1, 3], [0, 3, 4], [1, 0, 4]])
    torch.manual_seed(3)
    gcn = tg.EmbeddingGCN(At, X, edges, tg.create_matrix_M(4, 2), hidden_feat=[3, 2], condensed_W=True, use_Minv=False)
    out = gcn()
    assert tuple(out.shape) == (3, 2) and torch.isfinite(out).all()
    assert torch.count_nonzero(out[2]) == 0          # node 4 has no entry in slices 2 and 3 (band 2)
