"""(1) The BASELINE.json config shapes C1-C3 end to end against the oracle (SURVEY.md section 8, configs
1-3: Bitcoin-OTC-, chess- and SBM-shaped synthetic graphs, F 2 -> 6 -> 2, b = 20).
(2) Size-independent properties at the benchmark scale (N = 2M nodes, F = 128), where the oracle
cannot go: adjoint identities <K x, y> = <x, K^T y> for every forward/backward kernel pair,
linearity checksums and structural invariants of the sparse M-transform."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

TOL_OUT, TOL_GRAD = 1e-5, 1e-4


@pytest.fixture(scope="module")
def tg():
    import tmgcn_b200
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (the product has no CPU fallback)")
    tmgcn_b200._lib.load(build_if_missing=False)
    return tmgcn_b200


def relerr(got, ref):
    got, ref = got.detach().double().cpu(), torch.as_tensor(ref).detach().double().cpu()
    return ((got - ref).abs().max() / ref.abs().max()).item()


def sbm_coo(N, T, p_in, p_out, migrate, seed):
    """2-block dynamic SBM: `migrate` nodes change block every step (ref: SBM_our.py:101-109 uses
    dynamicgem's generator, which is not installed: regenerated here; data only, not arithmetic)."""
    rng = np.random.default_rng(seed)
    block = (np.arange(N) >= N // 2).astype(np.int64)
    ts, rs, cs = [], [], []
    for t in range(T):
        same = block[:, None] == block[None, :]
        prob = np.where(same, p_in, p_out)
        a = np.triu(rng.random((N, N)) < prob, 1)
        r, c = np.nonzero(a)
        ts.append(np.full(r.shape, t))
        rs.append(r)
        cs.append(c)
        mv = rng.choice(N, migrate, replace=False)
        block[mv] = 1 - block[mv]
    idx = np.stack([np.concatenate(ts), np.concatenate(rs), np.concatenate(cs)])
    return oracle.normalise_adjacency(idx, np.ones(idx.shape[1]), T, N)


@pytest.mark.parametrize("name,N,T,m,b", [("C1-bitcoin-otc", 5881, 95, 2580, 20), ("C2-chess", 7301, 79, 6500, 20),
                                          ("C3-sbm", 1000, 34, 0, 20)])
def test_config_shapes_vs_oracle(tg, name, N, T, m, b):
    from tmgcn_b200 import synth
    if name == "C3-sbm":
        idx, val = sbm_coo(N, T, 0.1, 0.01, 10, seed=3)
    else:
        i, v = synth.synth_coo(N, T, m, 0.9, seed=20261017)
        idx, val = i.numpy(), v.numpy()
    M = oracle.create_matrix_M(T, b)
    # (a) the sparse transform at the config's full size: bit-exact indices
    ref_idx, ref_val = oracle.func_MProduct(idx, val, (T, N, N), M.numpy(), no_diag=b)
    C = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), (T, N, N)).coalesce()
    Ct = tg.func_MProduct(C, M, no_diag=b)
    assert torch.equal(Ct._indices().cpu(), torch.from_numpy(ref_idx))
    np.testing.assert_allclose(Ct._values().cpu().numpy(), ref_val, rtol=1e-12, atol=0)
    # (b) the 2-layer model the experiment scripts build (experiment_bitcoin_our.py:107)
    At_ref = oracle.split_slices(ref_idx, ref_val, T, N)
    At = tg.split_slices(Ct.cpu())
    g = torch.Generator().manual_seed(7)
    X = torch.rand(T, N, 2, generator=g, dtype=torch.float64)
    E = 20000
    pick = torch.sort(torch.randint(0, ref_idx.shape[1], (E,), generator=g)).values
    edges = torch.from_numpy(ref_idx[:, pick.numpy()])
    torch.manual_seed(11)
    m_gpu = tg.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=True, use_Minv=False,
                             apply_M_twice=True, nonlin2="selu")
    m_ref = oracle.OracleGCN2(At_ref, X, edges, M, m_gpu.W1.detach().cpu(), m_gpu.W2.detach().cpu(),
                              m_gpu.U.detach().cpu(), apply_M_twice=True, nonlin2="selu", as_reference=False)
    dOut = torch.randn(E, 2, generator=g)
    out, out_r = m_gpu(), m_ref()
    assert relerr(out, out_r) <= TOL_OUT
    top2 = torch.topk(out_r.detach().double(), 2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 1e-4 * out_r.abs().max()
    assert torch.equal(out.argmax(1).cpu()[sure], out_r.argmax(1)[sure])
    out.backward(dOut.cuda())
    out_r.backward(dOut)
    for n in ("W1", "W2", "U"):
        assert relerr(getattr(m_gpu, n).grad, getattr(m_ref, n).grad) <= TOL_GRAD, n


def test_c4_shape_layer_vs_oracle(tg):
    """BASELINE.json configs[3] (Reddit shape: N = 55 863, b = 20, F = 128, experiment_reddit_our.py:31) with T cut
    to 44 slices -- what the CPU restatement does in seconds -- through the planned layer step: once as a single
    shard and once as the two time blocks of a 2-rank run chained on this GPU (halo = b-1 = 19 slices, the
    nearly-all-halo regime of C4 on 8 GPUs), both against the oracle's forward and backward."""
    from tmgcn_b200 import ops, synth
    from tmgcn_b200.layer_step import LayerStep
    N, T, m, b, F, C = 55_863, 44, 32_000, 20, 128, 2
    idx, val = synth.synth_coo(N, T, m, 0.9, seed=20261017)
    M = oracle.create_matrix_M(T, b)
    band = tg.Band(M)
    ref_idx, ref_val = oracle.func_MProduct(idx.numpy(), val.numpy(), (T, N, N), M.numpy(), no_diag=b)
    A = tg.SliceCSR.from_coo(idx, val, T, N)
    At = ops.mtransform_sparse(A, band)
    oi, ov = At.to_coo()
    assert torch.equal(oi.cpu(), torch.from_numpy(ref_idx))
    g = torch.Generator().manual_seed(5)
    H = torch.rand(T, N, F, generator=g)
    W = torch.randn(F, F, generator=g) / F ** 0.5
    U = torch.randn(2 * F, C, generator=g)
    E = 60_000
    pick = torch.sort(torch.randint(0, ref_idx.shape[1], (E,), generator=g)).values
    edges = torch.from_numpy(ref_idx[:, pick.numpy()])
    dOut = torch.randn(E, C, generator=g)
    out_r, dH_r, dW_r, dU_r = oracle.layer_fwd_bwd(oracle.split_slices(ref_idx, ref_val, T, N), H, M, W, U, edges,
                                                   dOut, "none", as_reference=False)
    Wd, Ud = W.cuda(), U.cuda()
    for mode in ("lowrank", "dense"):
        step = LayerStep(At, band, tg.EdgePlan(edges, N, T=T), F, F, C, "none", bwd_mode=mode)
        out = step.forward(H.cuda(), Wd, Ud)
        assert relerr(out, out_r) <= TOL_OUT, mode
        dH, dW, dU = step.backward(dOut.cuda(), Wd, Ud)
        assert relerr(dH, dH_r) <= TOL_GRAD and relerr(dW, dW_r) <= TOL_GRAD and relerr(dU, dU_r) <= TOL_GRAD, mode
        del step
    # two time blocks [0, 22) and [22, 44): the second holds a 19-slice halo of A and H; its halo gradient is
    # what a rank would send back to its predecessor
    t_cut, halo = 22, b - 1
    outs, dHs, dWs, dUs = [], torch.zeros(T, N, F), [], []
    for (t0, t1, h) in ((0, t_cut, 0), (t_cut, T, halo)):
        sel = (idx[0] >= t0 - h) & (idx[0] < t1)
        sub = idx[:, sel].clone()
        sub[0] -= t0 - h
        A_in = tg.SliceCSR.from_coo(sub, val[sel], t1 - t0 + h, N)
        At_b = ops.mtransform_sparse(A_in, band, t0, t1, h)
        esel = (edges[0] >= t0) & (edges[0] < t1)
        plan = tg.EdgePlan(edges[:, esel], N, t_offset=t0, T=t1 - t0)
        step = LayerStep(At_b, band, plan, F, F, C, "none", t0, t1, h, bwd_mode="dense")
        outs.append(step.forward(H[t0 - h:t1].cuda().contiguous(), Wd, Ud).cpu())
        dH, dW, dU = step.backward(dOut[esel].cuda().contiguous(), Wd, Ud)
        dHs[t0 - h:t1] += dH.cpu()
        dWs.append(dW.cpu().double())
        dUs.append(dU.cpu().double())
        del step, At_b, A_in
    assert relerr(torch.cat(outs), out_r) <= TOL_OUT
    assert relerr(dHs, dH_r) <= TOL_GRAD
    assert relerr(dWs[0] + dWs[1], dW_r) <= TOL_GRAD and relerr(dUs[0] + dUs[1], dU_r) <= TOL_GRAD


# --------------------------------------------------------------------------
# benchmark-scale properties (N = 2M, F = 128)
# --------------------------------------------------------------------------
N_BIG, F_BIG = 2_000_000, 128


def dot(a, b):
    return torch.dot(a.reshape(-1).double(), b.reshape(-1).double()).item()


@pytest.fixture(scope="module")
def big(tg):
    from tmgcn_b200 import ops, synth
    T, b = 4, 3
    A = synth.synth_csr(N_BIG, T, 2_000_000, 0.9, seed=5)
    band = tg.Band(tg.create_matrix_M(T, b))
    At = ops.mtransform_sparse(A, band)
    return dict(T=T, b=b, A=A, band=band, At=At)


def test_scale_sparse_transform_invariants(tg, big):
    from tmgcn_b200 import ops
    A, At, T, band = big["A"], big["At"], big["T"], big["band"]
    # every transformed row is strictly ascending in its columns (coalesce() order, ref: read_data.py:223)
    rid = At.row_ids()
    key = rid * N_BIG + At.col.to(torch.int64)
    assert bool((key[1:] > key[:-1]).all())
    # the union pattern only grows along the band; the first slice has nothing to merge with
    n_in, n_out = A.slice_nnz(), At.slice_nnz()
    assert bool((n_out >= n_in).all()) and int(n_out[0]) == int(n_in[0])
    assert int(n_out[1]) <= int(n_in[0] + n_in[1])
    # linearity checksum: sum(A~) = sum_t sum_s M[t, s] sum(A_s)
    s_in = torch.stack([A.val[A.rowptr[t * N_BIG]:A.rowptr[(t + 1) * N_BIG]].double().sum() for t in range(T)]).cpu()
    M = tg.create_matrix_M(T, big["b"])
    assert abs(At.val.double().sum().item() - float((M @ s_in).sum())) <= 1e-6 * float((M @ s_in).sum())
    # M = I is the identity, bit for bit; applying it twice is idempotent
    ident = tg.Band(torch.eye(T, dtype=torch.float64))
    same = ops.mtransform_sparse(A, ident)
    assert torch.equal(same.rowptr, A.rowptr) and torch.equal(same.col, A.col) and torch.equal(same.val, A.val)
    # transposing twice gives the matrix back, bit for bit
    tt = At.transpose()
    back = tg.SliceCSR(tt.T, tt.N, tt.rowptr, tt.col, tt.val).transpose()
    assert torch.equal(back.rowptr, At.rowptr) and torch.equal(back.col, At.col) and torch.equal(back.val, At.val)


def test_scale_adjoint_identities(tg, big):
    """<K x, y> = <x, K^T y> for stencil, SpMM, GEMM and readout at N = 2M, F = 128."""
    from tmgcn_b200 import ops
    T, band, At = big["T"], big["band"], big["At"]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(T, N_BIG, F_BIG, device="cuda", generator=g) - 0.5
    y = torch.rand(T, N_BIG, F_BIG, device="cuda", generator=g) - 0.5
    # stencil
    lhs, rhs = dot(ops.stencil_fwd(x, band), y), dot(x, ops.stencil_bwd(y, band))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # SpMM and its transposed CSR
    lhs, rhs = dot(ops.spmm_raw(At, x), y), dot(x, ops.spmm_raw(At.transpose(), y))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # skinny SpMM (the low-rank backward's 4-column factor)
    xs = torch.rand(T, N_BIG, 4, device="cuda", generator=g) - 0.5
    ys = torch.rand(T, N_BIG, 4, device="cuda", generator=g) - 0.5
    lhs, rhs = dot(ops.spmm_raw(At, xs), ys), dot(xs, ops.spmm_raw(At.transpose(), ys))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # GEMM: forward against true-fp32 cuBLAS on a slice, and <x W, y> = <x, y W^T> through the dP kernel
    W = torch.randn(F_BIG, F_BIG, device="cuda", generator=g) / F_BIG ** 0.5
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = x[0] @ W
    got = ops.gemm_fwd_raw(x[0].contiguous(), W)
    assert relerr(got, ref) <= TOL_OUT
    dp, dw = ops.gemm_bwd_raw(x[0].contiguous(), W, None, y[0].contiguous(), 0)
    lhs, rhs = dot(got, y[0]), dot(x[0], dp)
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # 2M-row reduction: the tensor-core partial sums are flushed every 2048 rows and added in rounded fp32
    assert relerr(dw, x[0].double().t() @ y[0].double()) <= 2e-5
    # readout: <R(Y), d> = <Y, R^T(d)> and linearity in U
    from tmgcn_b200 import synth
    E, Cc = 500_000, 2
    plan = tg.EdgePlan(synth.synth_edges(At, E, seed=2), N_BIG)
    U = torch.randn(2 * F_BIG, Cc, device="cuda", generator=g)
    d = torch.randn(E, Cc, device="cuda", generator=g)
    y2d = x.reshape(-1, F_BIG)
    out = ops.readout_fwd_raw(y2d, plan, U)
    dy, du = ops.readout_bwd_raw(y2d, plan, U, d)
    lhs = dot(out, d)
    assert abs(lhs - dot(y2d, dy)) <= 1e-5 * max(abs(lhs), 1.0)
    assert abs(lhs - dot(U, du)) <= 1e-5 * max(abs(lhs), 1.0)
