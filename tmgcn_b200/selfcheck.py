"""Multi-GPU self-check of the time-sharded, communication-overlapped layer step.

Every rank builds the same seeded graph, takes its (deliberately uneven) time block with the sparse halo
from its predecessor over NCCL, runs `LayerStep.forward/backward` with both halo implementations (NCCL
send/recv and the peer-memory fused stencil) and compares with the single-shard run of the whole tensor on
its own GPU.  `bench.py --gpus N` runs it before timing and prints the result as `parity_multi_gpu`;
`tests/multi_gpu_check.py` is the stand-alone driver.

This compares the product with itself (sharded against unsharded); the comparison with the reference's
arithmetic is the job of the single-GPU parity tests (tests/test_gpu_parity.py).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.distributed as dist

from . import ops, sharding, synth
from .layer_step import LayerStep
from .modules import create_matrix_M
from .ops import Band, EdgePlan, SliceCSR


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    d = b.double().abs().max().item()
    return (a.double() - b.double()).abs().max().item() / (d if d > 0 else 1.0)


def multi_gpu_parity(rank: int, world: int, dev, acts: Sequence[str] = ("relu", "none"), N: int = 20000,
                     F: int = 128, C: int = 2, b: int = 5, slices_per_rank: int = 12, peer_storage=None,
                     try_peer: bool = True) -> Dict:
    """-> {"ok": bool, "bounds": [...], act: {...errors...}}; collective over the default process group."""
    T = slices_per_rank * world + 3                       # + 3: the blocks cannot all have the same length
    idx, val = synth.synth_coo(N, T, 3 * N, 0.85, seed=11, device="cpu")
    M = create_matrix_M(T, b)
    band = Band(M)
    g = torch.Generator().manual_seed(3)
    H = torch.rand(T, N, F, generator=g)
    W = (torch.randn(F, F, generator=g) / F ** 0.5).to(dev)
    U = torch.randn(2 * F, C, generator=g).to(dev)
    full_At = ops.mtransform_sparse(SliceCSR.from_coo(idx, val, T, N), band)
    E = 2 * N
    edges = synth.synth_edges(full_At, E, seed=5)          # same on every rank (same seed, same tensor)
    dOut = torch.randn(E, C, generator=g).to(dev)

    w_est = [float(x) for x in full_At.slice_nnz().tolist()]
    bounds = sharding.balanced_bounds(w_est, world)
    t0, t1 = bounds[rank]
    Tl = t1 - t0
    halo = min(b - 1, t0) if rank > 0 else 0
    sel = (idx[0] >= t0) & (idx[0] < t1)
    own_idx = idx[:, sel].clone()
    own_idx[0] -= t0
    A_own = SliceCSR.from_coo(own_idx, val[sel], Tl, N)
    A_in = sharding.exchange_sparse_halo(A_own, b - 1, rank, world)
    At = ops.mtransform_sparse(A_in, band, t0, t1, halo)
    esel = (edges[0] >= t0) & (edges[0] < t1)
    plan = EdgePlan(edges[:, esel], N, t_offset=t0, T=Tl)
    dOut_l = dOut[esel].contiguous()
    comm = sharding.ShardComm(b - 1, rank, world, dev, T_own=Tl)
    peer = None
    if try_peer:
        try:
            peer = sharding.PeerHalo(Tl, N, F, b - 1, rank, world, dev, storage=peer_storage)
        except Exception:                                  # symmetric memory is optional on the box
            peer = None
        okp = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(okp, op=dist.ReduceOp.MIN)
        if int(okp.item()) == 0:
            peer = None
    ref_plan = EdgePlan(edges, N, T=T)
    result: Dict = {"ok": True, "world": world, "bounds": [list(x) for x in bounds], "T": T, "N": N, "F": F, "C": C,
                    "b": b, "halo_modes": ["nccl"] + (["peer"] if peer is not None else [])}
    for act in acts:
        step = LayerStep(At, band, plan, F, F, C, act, t0, t1, halo)
        Hl = torch.zeros(halo + Tl, N, F, device=dev)
        Hl[halo:] = H[t0:t1].to(dev)                       # the halo slices arrive over NCCL
        for _ in range(2):                                 # twice: buffers and events are re-used
            out = step.forward(Hl, W, U, comm).clone()
            dH, dW, dU = step.backward(dOut_l, W, U, comm)
            dH, dW, dU = dH[halo:].clone(), dW.clone(), dU.clone()
            torch.cuda.synchronize()
        halo_ok = bool(torch.equal(Hl[:halo].cpu(), H[t0 - halo:t0]))
        peer_equal = None
        if peer is not None:
            peer.H.copy_(H[t0:t1].to(dev))
            for _ in range(2):
                out_p = step.forward(peer.H, W, U, comm, peer).clone()
                dH_p, dW_p, dU_p = step.backward(dOut_l, W, U, comm)
                torch.cuda.synchronize()
            peer_equal = bool(torch.equal(out_p, out) and torch.equal(dW_p, dW) and torch.equal(dU_p, dU)
                              and torch.equal(dH_p[halo:], dH))
        del step
        def reference():
            r_ = LayerStep(full_At, band, ref_plan, F, F, C, act, bwd_mode="dense")
            o_ = r_.forward(H.to(dev), W, U).clone()
            dh_, dw_, du_ = r_.backward(dOut, W, U)
            res_ = (o_, dh_[t0:t1].clone(), dw_.clone(), du_.clone())
            torch.cuda.synchronize()
            del r_
            return res_
        out_r, dH_r, dW_r, dU_r = reference()
        errs = {"out": _rel(out, out_r[esel]), "dH": _rel(dH, dH_r), "dW": _rel(dW, dW_r), "dU": _rel(dU, dU_r)}
        first_attempt = None
        if errs["out"] > 1e-5 or max(errs["dH"], errs["dW"], errs["dU"]) > 1e-4:
            # do not hide it: recompute the single-GPU reference once and say which side moved
            out2, dH2, dW2, dU2 = reference()
            first_attempt = {"rel_err": dict(errs),
                             "reference_changed_on_recompute": {"out": _rel(out2, out_r), "dH": _rel(dH2, dH_r),
                                                                "dW": _rel(dW2, dW_r), "dU": _rel(dU2, dU_r)}}
            out_r, dH_r, dW_r, dU_r = out2, dH2, dW2, dU2
            errs = {"out": _rel(out, out_r[esel]), "dH": _rel(dH, dH_r), "dW": _rel(dW, dW_r), "dU": _rel(dU, dU_r)}
        bit = bool(torch.equal(out, out_r[esel]))
        tri = None
        if act in (None, "none"):
            # triangulate the two backward formulations on the unsharded tensor as well: a disagreement between
            # them is a kernel problem, a disagreement of the sharded run with both is a sharding problem
            ref = LayerStep(full_At, band, ref_plan, F, F, C, act, bwd_mode="lowrank")
            ref.forward(H.to(dev), W, U)
            _, _, dU_l = ref.backward(dOut, W, U)
            tri = {"single_lowrank_vs_single_dense": _rel(dU_l, dU_r), "sharded_vs_single_lowrank": _rel(dU, dU_l)}
            del ref
        torch.cuda.empty_cache()
        red = torch.tensor([errs["out"], errs["dH"], errs["dW"], errs["dU"], 0.0 if bit else 1.0,
                            0.0 if halo_ok else 1.0, 0.0 if peer_equal in (None, True) else 1.0],
                           device=dev, dtype=torch.float64)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        r = [float(x) for x in red.tolist()]
        entry = {"backward": "lowrank" if act in (None, "none") else "dense",
                 "rel_err_out": r[0], "rel_err_dH": r[1], "rel_err_dW": r[2], "rel_err_dU": r[3],
                 "logits_bit_equal_to_single_gpu": r[4] == 0.0, "nccl_halo_content_exact": r[5] == 0.0,
                 "peer_halo_bit_equal_to_nccl": (None if peer is None else r[6] == 0.0)}
        if tri is not None:
            entry["dU_triangulation"] = tri
        gathered = [None] * world
        dist.all_gather_object(gathered, first_attempt)
        if any(x is not None for x in gathered):
            # a rank's first comparison failed and it recomputed its single-GPU reference: reported, not hidden
            entry["first_attempt_failed"] = {f"rank{i}": x for i, x in enumerate(gathered) if x is not None}
        entry["ok"] = (r[0] <= 1e-5 and max(r[1], r[2], r[3]) <= 1e-4 and r[5] == 0.0 and r[6] == 0.0)
        result[str(act)] = entry
        result["ok"] = result["ok"] and entry["ok"]
    # use_Minv across ranks: the pipelined cross-rank substitution (forward and adjoint) against the
    # single-GPU solve of the whole tensor
    Zs = torch.randn(T, N, 8, generator=g).to(dev)
    full_fwd = ops._Solve.apply(Zs, band)
    gz = torch.empty_like(Zs)
    from . import _lib
    _lib.check(_lib.load().tmgcn_mtransform_dense_solve_bwd(ops._p(Zs), ops._p(gz), T, Zs[0].numel(),
                                                            ops._p(band.device_weights(0, T, torch.float32)), b,
                                                            ops._stream()))
    e_f = _rel(sharding.solve_pipelined(Zs[t0:t1].contiguous(), band, t0, t1, rank, world), full_fwd[t0:t1])
    e_b = _rel(sharding.solve_pipelined(Zs[t0:t1].contiguous(), band, t0, t1, rank, world, transposed=True), gz[t0:t1])
    red = torch.tensor([e_f, e_b], device=dev, dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    e_f, e_b = (float(x) for x in red.tolist())
    result["use_minv_pipelined_solve"] = {"rel_err_fwd": e_f, "rel_err_adjoint": e_b, "ok": max(e_f, e_b) <= 1e-5}
    result["ok"] = result["ok"] and result["use_minv_pipelined_solve"]["ok"]
    result["tolerance"] = {"out": 1e-5, "grads": 1e-4}
    return result
