"""Graph preparation on the device -- the step in front of the sparse M-transform
("next" row 1; ref: TensorGCN-master/read_data.py:88-188).

Same function names and argument meaning as the reference's module-level helpers; inputs and
outputs are coalesced sparse COO tensors (T, N, N), values fp64, indices in (t, i, j) order.
The `*_csr` variants keep everything in the CSR-of-slices layout for a device-resident pipeline.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .ops import Band, SliceCSR, _p, _stream


def _to_csr(A: torch.Tensor, dtype=torch.float64) -> SliceCSR:
    A = A if A.is_coalesced() else A.coalesce()
    T, N, N2 = A.shape
    if N != N2:
        raise ValueError("expected a T x N x N tensor")
    return SliceCSR.from_coo(A._indices(), A._values(), T, N, dtype=dtype)


def _to_coo(C: SliceCSR) -> torch.Tensor:
    idx, val = C.to_coo()
    return torch.sparse_coo_tensor(idx, val, (C.T, C.N, C.N), is_coalesced=True)


csr_axpby = ops.csr_axpby


def make_symmetric_csr(A: SliceCSR) -> SliceCSR:
    """(A_t + A_t^T) / 2 per slice (ref: read_data.py:88-109)."""
    return csr_axpby(A, A.transpose(), 0.5, 0.5)


def edge_life_csr(A: SliceCSR, edge_life_window: int) -> SliceCSR:
    """A_new[t] = sum of A[s], s in [max(0, t-w+1), t] (ref: read_data.py:116-125): a ones-band M-transform."""
    M = torch.zeros(A.T, A.T, dtype=torch.float64)
    for i in range(min(edge_life_window, A.T)):
        M.diagonal(-i).fill_(1.0)
    return ops.mtransform_sparse(A, Band(M))


def laplacian_transformation_csr(B: SliceCSR) -> SliceCSR:
    """D^-1/2 (B + I) D^-1/2 per slice, D = row sums of B + I (ref: read_data.py:130-164)."""
    lib = _lib.load()
    T, N, dev = B.T, B.N, B.rowptr.device
    n_rows = T * N
    eye = SliceCSR(T, N, torch.arange(n_rows + 1, dtype=torch.int64, device=dev),
                   (torch.arange(n_rows, device=dev) % N).to(torch.int32),
                   torch.ones(n_rows, dtype=B.val.dtype, device=dev))
    C = csr_axpby(B, eye, 1.0, 1.0)
    f64 = 1 if C.val.dtype == torch.float64 else 0
    deg = torch.empty(n_rows, dtype=torch.float64, device=dev)
    _lib.check(lib.tmgcn_csr_row_sums(_p(C.rowptr), _p(C.val), n_rows, _p(deg), f64, _stream()))
    _lib.check(lib.tmgcn_csr_scale_sym(_p(C.rowptr), _p(C.col), _p(C.val), T, N, _p(deg), f64, _stream()))
    return C


def create_sparse_csr(A: SliceCSR, start: int, end: int) -> SliceCSR:
    """time window [start, end) re-based to 0 (ref: read_data.py:174-183)."""
    return A.time_window(start, end)


# ---- the reference's call surface (sparse COO in, sparse COO out) -------------------------------
def func_make_symmetric(sparse_tensor: torch.Tensor, N: int, TT: int) -> torch.Tensor:
    return _to_coo(make_symmetric_csr(_to_csr(sparse_tensor)))


def func_edge_life(A: torch.Tensor, N: int, TT: int, edge_life_window: int = 10) -> torch.Tensor:
    return _to_coo(edge_life_csr(_to_csr(A), edge_life_window))


def func_laplacian_transformation(B: torch.Tensor, N: int, TT: int) -> torch.Tensor:
    return _to_coo(laplacian_transformation_csr(_to_csr(B)))


def func_create_sparse(A: torch.Tensor, N: int, TTT: int, T: int, start: int, end: int) -> torch.Tensor:
    assert (end - start) == T  # ref: read_data.py:175
    return _to_coo(create_sparse_csr(_to_csr(A), start, end))
