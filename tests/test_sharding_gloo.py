"""world_size-2/3 gloo tests of the time-sharding host logic (halo exchange both ways,
sparse halo, gradient all-reduce) on CPU tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, N, F, b, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tmgcn_b200 import sharding, synth
        from tmgcn_b200.ops import SliceCSR
        torch.manual_seed(0)
        M = oracle.create_matrix_M(T, b)
        X = torch.rand(T, N, F, dtype=torch.float64)
        G = torch.randn(T, N, F, dtype=torch.float64)
        ref = (M @ X.reshape(T, -1)).reshape(X.shape)
        ref_g = (M.T @ G.reshape(T, -1)).reshape(X.shape)
        t0, t1 = sharding.shard_bounds(T, world, rank)
        Tl = t1 - t0
        halo = min(b - 1, t0) if rank > 0 else 0
        hx = sharding.DenseHalo(N * F, b - 1, rank, world)
        # ---- forward halo: own slices known, halo slices received
        H = torch.zeros(halo + Tl, N, F, dtype=torch.float64)
        H[halo:] = X[t0:t1]
        hx.forward(H, Tl, halo)
        assert torch.equal(H[:halo], X[t0 - halo:t0])
        Ml = M[t0:t1, t0 - halo:t1]
        out = (Ml @ H.reshape(halo + Tl, -1)).reshape(Tl, N, F)
        assert torch.allclose(out, ref[t0:t1], rtol=1e-13, atol=1e-13)
        # ---- backward halo: partial sums for the predecessor travel back and are added
        dH = (Ml.T @ G[t0:t1].reshape(Tl, -1)).reshape(halo + Tl, N, F)
        scratch = torch.empty((b - 1) * N * F + 5, dtype=torch.float64)
        hx.backward(dH, Tl, halo, scratch=scratch)
        assert torch.allclose(dH[halo:], ref_g[t0:t1], rtol=1e-13, atol=1e-13)
        # ---- sparse halo (once per dataset)
        idx, val = synth.synth_coo(N, T, 3 * N, 0.7, seed=5)
        sel = (idx[0] >= t0) & (idx[0] < t1)
        own_idx = idx[:, sel].clone()
        own_idx[0] -= t0
        flat = own_idx[0] * N + own_idx[1]
        rowptr = torch.searchsorted(flat, torch.arange(Tl * N + 1))
        A_own = SliceCSR(Tl, N, rowptr, own_idx[2].to(torch.int32), val[sel].float())
        A_in = sharding.exchange_sparse_halo(A_own, b - 1, rank, world)
        h_in = A_in.T - Tl
        assert h_in == (min(b - 1, t0 - sharding.shard_bounds(T, world, rank - 1)[0]) if rank > 0 else 0)
        sel2 = (idx[0] >= t0 - h_in) & (idx[0] < t1)
        exp_idx = idx[:, sel2].clone()
        exp_idx[0] -= t0 - h_in
        exp_rowptr = torch.searchsorted(exp_idx[0] * N + exp_idx[1], torch.arange((Tl + h_in) * N + 1))
        assert torch.equal(A_in.rowptr, exp_rowptr)
        assert torch.equal(A_in.col, exp_idx[2].to(torch.int32))
        assert torch.equal(A_in.val, val[sel2].float())
        # ---- gradient all-reduce
        dW, dU = torch.full((3, 4), float(rank + 1)), torch.full((2,), 10.0 * (rank + 1))
        sharding.allreduce_grads([dW, dU])
        tot = world * (world + 1) / 2
        assert torch.all(dW == tot) and torch.all(dU == 10 * tot)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _cpu_local_solve(src, dst, halo, T, h, n, ld, ld_halo, w, b, transposed):
    """torch restatement of one column chunk of the banded substitution (the device kernel's contract:
    include/tmgcn.h, tmgcn_mtransform_dense_solve_part); fp64 recurrence, rows are views of width >= n."""
    w = w.double()
    out = torch.zeros(T, n, dtype=torch.float64)
    z = src[:, :n].double()
    hal = halo[:, :n].double() if h > 0 else None
    if not transposed:
        for t in range(T):
            acc = z[t].clone()
            for i in range(1, b):
                prev = out[t - i] if t - i >= 0 else (hal[h + (t - i)] if hal is not None and h + (t - i) >= 0 else None)
                if prev is not None:
                    acc -= w[t, i] * prev
            out[t] = acc / w[t, 0]
    else:
        for t in range(T - 1, -1, -1):
            acc = z[t].clone()
            for i in range(1, b):
                if t + i < T:
                    acc -= w[t + i, i] * out[t + i]
                elif hal is not None and t + i - T < h and t + i < w.shape[0]:
                    acc -= w[t + i, i] * hal[t + i - T]
            out[t] = acc / w[t, 0]
    dst[:, :n] = out.to(dst.dtype)


def _solve_worker(rank, world, port, T, b, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tmgcn_b200 import sharding
        from tmgcn_b200.ops import Band
        torch.manual_seed(1)
        M = oracle.create_matrix_M(T, b, normalize=True)
        band = Band(M)
        Z = torch.randn(T, 7, 3, dtype=torch.float64)
        Minv = torch.linalg.inv(M)
        ref = (Minv @ Z.reshape(T, -1)).reshape(Z.shape)
        ref_t = (Minv.T @ Z.reshape(T, -1)).reshape(Z.shape)
        blocks = sharding.balanced_bounds([1.0 + 0.3 * (t % 5) for t in range(T)], world)   # uneven blocks
        t0, t1 = blocks[rank]
        for transposed, want in ((False, ref), (True, ref_t)):
            y = sharding.solve_pipelined(Z[t0:t1].contiguous(), band, t0, t1, rank, world, transposed=transposed,
                                         chunks=4, local_solve=_cpu_local_solve)
            err = ((y - want[t0:t1]).abs().max() / want.abs().max()).item()
            assert err < 1e-6, (transposed, err)                     # band weights travel as fp32
        ret[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T,b", [(2, 14, 4), (3, 30, 6)])
def test_pipelined_solve_gloo(world, T, b):
    """use_Minv across ranks: the cross-rank scan, forward and adjoint, against inv(M)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_solve_worker, args=(world, _free_port(), T, b, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


def test_balanced_bounds():
    from tmgcn_b200 import sharding
    w = sharding.slice_weight_estimate(256, 10, 2_000_000, 10_000_000, 0.9)
    blocks = sharding.balanced_bounds(w, 8)
    assert blocks[0][0] == 0 and blocks[-1][1] == 256
    assert all(a[1] == b_[0] for a, b_ in zip(blocks[:-1], blocks[1:]))
    loads = [sum(w[a:b_]) for a, b_ in blocks]
    assert max(loads) / min(loads) < 1.05                            # equal slice counts would give 1.20
    assert blocks[0][1] - blocks[0][0] > blocks[1][1] - blocks[1][0]  # rank 0's truncated windows are cheaper
    assert sharding.balanced_bounds([1.0] * 7, 7) == [(i, i + 1) for i in range(7)]
    with pytest.raises(ValueError):
        sharding.balanced_bounds([1.0] * 3, 4)


@pytest.mark.parametrize("world,T,b", [(2, 12, 4), (3, 14, 3), (2, 40, 20), (4, 36, 10)])
def test_time_sharding_gloo(world, T, b):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), T, 6, 5, b, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


def test_shard_bounds():
    from tmgcn_b200 import sharding
    for T, w in [(256, 8), (10, 3), (7, 7), (178, 8)]:
        blocks = [sharding.shard_bounds(T, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == T
        assert all(a[1] == b_[0] for a, b_ in zip(blocks[:-1], blocks[1:]))
        sizes = [b_ - a for a, b_ in blocks]
        assert max(sizes) - min(sizes) <= 1
