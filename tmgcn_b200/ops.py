"""Tensor-facing wrappers over the C ABI (include/tmgcn.h) and their autograd wiring.

PyTorch is used for device memory, streams and autograd bookkeeping only; all
arithmetic on the hot path happens in libtmgcn_b200.so.  There is no CPU
fallback: tensors are moved to the current CUDA device and the library raises if
it cannot run.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib

ACT = {None: 0, "none": 0, "relu": 1, "leaky": 2, "selu": 3}


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tmgcn_b200 needs a CUDA device (no CPU fallback exists)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> C.c_void_p:
    """raw device pointer of a contiguous CUDA tensor; `dtype` = what the kernel will read it as (the
    library takes untyped pointers, so a tensor of another dtype would be silently reinterpreted)."""
    if t is None:
        return C.c_void_p(0)
    if not (t.is_cuda and t.is_contiguous()):
        raise ValueError("tmgcn: expected a contiguous CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"tmgcn: expected a {dtype} tensor, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 tensor on the current CUDA device (no copy when it already is one): every public
    entry point funnels caller tensors through this, e.g. the reference's fp64 X (ehf:204 feeds .double())"""
    if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
        return t
    return t.to(device=_dev(), dtype=torch.float32).contiguous()


_F, _L, _I = torch.float32, torch.int64, torch.int32


def _ws(nbytes: int) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=_dev())


# --------------------------------------------------------------------------
# integer helpers
# --------------------------------------------------------------------------
def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    """int64 counts[n] -> int64 out[n+1] (out[n] = total)."""
    lib = _lib.load()
    counts = counts.to(device=_dev(), dtype=torch.int64).contiguous()
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    ws = _ws(lib.tmgcn_scan_ws_bytes(n))
    _lib.check(lib.tmgcn_exclusive_scan_i64(_p(counts), _p(out), n, _p(ws), _stream()))
    return out


class SliceCSR:
    """T x N x N sparse tensor as one CSR over T*N rows (see include/tmgcn.h)."""

    def __init__(self, T: int, N: int, rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor):
        self.T, self.N = int(T), int(N)
        self.rowptr, self.col, self.val = rowptr, col, val
        self._t: Optional["SliceCSR"] = None

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def slice_nnz(self) -> torch.Tensor:
        rp = self.rowptr[:: self.N] if self.N > 0 else self.rowptr
        return rp[1:] - rp[:-1]

    @staticmethod
    def from_coo(idx: torch.Tensor, val: torch.Tensor, T: int, N: int, dtype=torch.float32) -> "SliceCSR":
        """idx (3, nnz) int64 in coalesced (t, i, j) order
        (replaces the per-slice masking of ref: ehf:561-572)."""
        lib = _lib.load()
        dev = _dev()
        idx = idx.to(device=dev, dtype=torch.int64)
        if idx.dim() != 2 or idx.shape[0] != 3 or val.numel() != idx.shape[1]:
            raise ValueError("from_coo: idx must be (3, nnz) with one value per column")
        flat = (idx[0] * N + idx[1]).contiguous()
        if flat.numel():
            # validated once here: the row-pointer kernel writes rowptr[r] for every r up to rows[k], so an
            # out-of-range or unsorted entry would be an out-of-bounds device write / uninitialised rowptr
            lo = torch.stack([idx[0].min(), idx[1].min(), idx[2].min()])
            hi = torch.stack([idx[0].max() - T, idx[1].max() - N, idx[2].max() - N])
            if bool((lo < 0).any()) or bool((hi >= 0).any()):
                raise ValueError(f"from_coo: an index is outside the {T} x {N} x {N} tensor")
            key = flat * N + idx[2]
            if bool((key[1:] <= key[:-1]).any()):
                raise ValueError("from_coo: entries must be coalesced (strictly ascending (t, i, j) order)")
            del key
        col = idx[2].to(torch.int32).contiguous()
        val = val.to(device=dev, dtype=dtype).contiguous()
        rowptr = torch.empty(T * N + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.tmgcn_rowptr_from_sorted_rows(_p(flat), flat.numel(), T * N, _p(rowptr), _stream()))
        return SliceCSR(T, N, rowptr, col, val)

    @staticmethod
    def from_slice_list(slices: Sequence[torch.Tensor], N: int, dtype=torch.float32) -> "SliceCSR":
        """Python list of T 2-D sparse COO matrices (the reference's `At` argument,
        ref: experiment_bitcoin_our.py:53-56) -> device CSR-of-slices.  The list lives on the host in the
        reference's scripts, and this conversion is most of a fresh-input call (ehf:212-215), so it moves as
        little as it can: (row, col) as int32 and the values in the target dtype, 12 bytes per entry instead of
        the 32 of an int64 (t, i, j) + fp64 COO; the time index is rebuilt on the device from the slice lengths."""
        lib = _lib.load()
        dev = _dev()
        T = len(slices)
        ijs, vs, lens = [], [], []
        for t, A in enumerate(slices):
            if A.layout != torch.sparse_coo:
                raise TypeError("every slice must be a sparse COO tensor")
            A = A if A.is_coalesced() else A.coalesce()
            ijs.append(A._indices())
            vs.append(A._values())
            lens.append(A._nnz())
        if T == 0 or sum(lens) == 0:
            return SliceCSR.from_coo(torch.zeros(3, 0, dtype=torch.int64), torch.zeros(0, dtype=torch.float64), T, N, dtype)
        ij = torch.cat(ijs, dim=1)
        if N >= 2 ** 31:
            raise ValueError("N must fit int32")
        ij = ij.to(torch.int32).to(dev, non_blocking=True)                   # (2, nnz) int32
        val = torch.cat(vs).to(dtype).to(dev, non_blocking=True).contiguous()
        lens_d = torch.tensor(lens, dtype=torch.int64).to(dev)
        # range and order checks (what from_coo does), on the device
        lo, hi = ij.min(), ij.max()
        if int(lo) < 0 or int(hi) >= N:
            raise ValueError(f"a slice has an index outside [0, {N})")
        t_idx = torch.repeat_interleave(torch.arange(T, device=dev, dtype=torch.int64), lens_d)
        flat = (t_idx * N + ij[0]).contiguous()
        key = flat * N + ij[1]
        if bool((key[1:] <= key[:-1]).any()):
            raise ValueError("from_slice_list: slices must be coalesced (strictly ascending (i, j) order)")
        del key, t_idx
        rowptr = torch.empty(T * N + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.tmgcn_rowptr_from_sorted_rows(_p(flat), flat.numel(), T * N, _p(rowptr), _stream()))
        return SliceCSR(T, N, rowptr, ij[1].contiguous(), val)

    def time_window(self, start: int, end: int) -> "SliceCSR":
        """slices [start, end) re-based to 0 (ref: func_create_sparse, read_data.py:174-183)"""
        N = self.N
        lo, hi = int(self.rowptr[start * N].item()), int(self.rowptr[end * N].item())
        return SliceCSR(end - start, N, (self.rowptr[start * N:end * N + 1] - lo).contiguous(),
                        self.col[lo:hi].contiguous(), self.val[lo:hi].contiguous())

    def concat(self, other: "SliceCSR") -> "SliceCSR":
        """[self | other] along time"""
        assert self.N == other.N and self.val.dtype == other.val.dtype
        rowptr = torch.cat([self.rowptr[:-1], other.rowptr + self.rowptr[-1]])
        return SliceCSR(self.T + other.T, self.N, rowptr, torch.cat([self.col, other.col]),
                        torch.cat([self.val, other.val]))

    def row_ids(self) -> torch.Tensor:
        counts = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.T * self.N, device=self.rowptr.device), counts)

    def to_coo(self):
        """-> idx (3, nnz) int64 in coalesced order, val."""
        r = self.row_ids()
        idx = torch.stack([r // self.N, r % self.N, self.col.to(torch.int64)])
        return idx, self.val

    def transpose(self) -> "SliceCSR":
        """Per-slice transpose (cached) for the backward SpMM."""
        if self._t is not None:
            return self._t
        f64 = 1 if self.val.dtype == torch.float64 else 0
        lib = _lib.load()
        dev = self.rowptr.device
        n_rows = self.T * self.N
        counts = torch.empty(n_rows, dtype=torch.int64, device=dev)
        _lib.check(lib.tmgcn_csr_transpose_plan(_p(self.rowptr), _p(self.col), self.T, self.N, _p(counts), _stream()))
        t_rowptr = exclusive_scan(counts)
        del counts
        t_col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        t_val = torch.empty(self.nnz, dtype=self.val.dtype, device=dev)
        ws = _ws(lib.tmgcn_csr_transpose_ws_bytes(n_rows, self.nnz, f64))
        _lib.check(lib.tmgcn_csr_transpose_run(_p(self.rowptr), _p(self.col), _p(self.val), self.T, self.N,
                                               _p(t_rowptr), _p(t_col), _p(t_val), f64, _p(ws), _stream()))
        self._t = SliceCSR(self.T, self.N, t_rowptr, t_col, t_val)
        self._t._t = self
        return self._t


# --------------------------------------------------------------------------
# M as a band
# --------------------------------------------------------------------------
class Band:
    """Banded lower-triangular M (ref: SBM_our.py:88-96, read_data.py:56-62):
    w[t, i] = M[t, t-i].  `rows` selects the output slices a rank owns."""

    MAX_KERNEL_BAND = 32      # widest band one kernel launch handles (register ring / cursor count)

    def __init__(self, M: torch.Tensor):
        M = torch.as_tensor(M).detach().to("cpu", torch.float64)
        if M.dim() != 2 or M.shape[0] != M.shape[1]:
            raise ValueError("M must be a square matrix")
        if torch.count_nonzero(torch.triu(M, 1)) != 0:
            raise NotImplementedError("M must be lower triangular (banded); dense M / inv(M) is out of scope")
        T = M.shape[0]
        nz = torch.nonzero(M)
        b = int((nz[:, 0] - nz[:, 1]).max()) + 1 if nz.numel() else 1
        w = torch.zeros(T, b, dtype=torch.float64)
        for i in range(b):
            w[i:, i] = torch.diagonal(M, -i)
        self.T, self.b, self.w = T, b, w
        self._dev = {}

    @classmethod
    def _from_weights(cls, w: torch.Tensor) -> "Band":
        self = cls.__new__(cls)
        self.T, self.b, self.w, self._dev = int(w.shape[0]), int(w.shape[1]), w.contiguous(), {}
        return self

    def chunks(self):
        """A band wider than one kernel handles, as lag blocks: [(o, sub)] with sub.w[t', i] = w[t' + o, o + i],
        so that M x_3 X = sum over blocks of (sub x_3 X[:T-o]) placed at output slices [o, T)."""
        out = []
        for o in range(0, self.b, self.MAX_KERNEL_BAND):
            if o < self.T:
                out.append((o, Band._from_weights(self.w[o:, o:min(self.b, o + self.MAX_KERNEL_BAND)])))
        return out

    def _whole_only(self, t0, t1, halo, what):
        if self.b > self.MAX_KERNEL_BAND and not (t0 == 0 and t1 == self.T and halo == 0):
            raise NotImplementedError(f"{what}: bands wider than {self.MAX_KERNEL_BAND} are only supported unsharded")

    def device_weights(self, t0: int, t1: int, dtype) -> torch.Tensor:
        key = (t0, t1, dtype, torch.cuda.current_device())
        if key not in self._dev:
            self._dev[key] = self.w[t0:t1].to(device=_dev(), dtype=dtype).contiguous()
        return self._dev[key]


def mtransform_sparse(csr: SliceCSR, band: Band, t0: int = 0, t1: Optional[int] = None, halo: int = 0) -> SliceCSR:
    """A~ = A x_3 M on CSR-of-slices (ref: func_MProduct, read_data.py:204-223).
    `csr` holds input slices [t0 - halo, t1); the result holds output slices [t0, t1)."""
    lib = _lib.load()
    t1 = band.T if t1 is None else t1
    T_out = t1 - t0
    if csr.T != T_out + halo:
        raise ValueError(f"input has {csr.T} slices, expected halo + T_out = {halo + T_out}")
    if band.b > Band.MAX_KERNEL_BAND:
        # wide band: one merge per block of 32 lags on the time-shifted input, partial tensors added by the
        # sorted 2-way row merge (same union pattern and order; the fp64 sums are grouped per block)
        band._whole_only(t0, t1, halo, "mtransform_sparse")
        out = None
        for o, sub in band.chunks():
            part = mtransform_sparse(csr.time_window(0, csr.T - o), sub)
            if out is None:
                out = part
            else:
                tail = csr_axpby(out.time_window(o, out.T), part, 1.0, 1.0)
                out = out.time_window(0, o).concat(tail)
        return out
    f64 = csr.val.dtype == torch.float64
    w = band.device_weights(t0, t1, torch.float64)
    N = csr.N
    dev = csr.rowptr.device
    counts = torch.empty(T_out * N, dtype=torch.int64, device=dev)
    # workspace of the union-list variant (0 bytes = not applicable: the fill pass merges again)
    ws_bytes = 0 if f64 else int(lib.tmgcn_mtransform_sparse_ws_bytes(T_out, halo, N, band.b, csr.nnz))   # fp32 layout only
    ws = None
    if ws_bytes:
        try:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        except torch.cuda.OutOfMemoryError:
            ws_bytes = 0
    _lib.check(lib.tmgcn_mtransform_sparse_plan_ws(_p(csr.rowptr), _p(csr.col), T_out, halo, N, _p(w), band.b,
                                                   _p(counts), _p(ws), ws_bytes, _stream()))
    rowptr = exclusive_scan(counts)
    del counts
    nnz = int(rowptr[-1].item())
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(nnz, dtype=csr.val.dtype, device=dev)
    _lib.check(lib.tmgcn_mtransform_sparse_run_ws(_p(csr.rowptr), _p(csr.col), _p(csr.val), T_out, halo, N, _p(w),
                                                  band.b, _p(rowptr), _p(col), _p(val), 1 if f64 else 0,
                                                  _p(ws), ws_bytes, _stream()))
    return SliceCSR(T_out, N, rowptr, col, val)


def csr_axpby(A: SliceCSR, B: SliceCSR, alpha: float, beta: float) -> SliceCSR:
    """alpha*A + beta*B (sorted 2-way row merge; union pattern, explicit zeros kept)."""
    assert A.T == B.T and A.N == B.N and A.val.dtype == B.val.dtype
    lib = _lib.load()
    n_rows = A.T * A.N
    dev = A.rowptr.device
    counts = torch.empty(n_rows, dtype=torch.int64, device=dev)
    _lib.check(lib.tmgcn_csr_axpby_plan(_p(A.rowptr), _p(A.col), _p(B.rowptr), _p(B.col), n_rows, _p(counts), _stream()))
    rowptr = exclusive_scan(counts)
    nnz = int(rowptr[-1].item())
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(nnz, dtype=A.val.dtype, device=dev)
    _lib.check(lib.tmgcn_csr_axpby_run(_p(A.rowptr), _p(A.col), _p(A.val), _p(B.rowptr), _p(B.col), _p(B.val),
                                       float(alpha), float(beta), n_rows, _p(rowptr), _p(col), _p(val),
                                       1 if A.val.dtype == torch.float64 else 0, _stream()))
    return SliceCSR(A.T, A.N, rowptr, col, val)


# --------------------------------------------------------------------------
# raw (non-autograd) kernels
# --------------------------------------------------------------------------
def stencil_fwd(x: torch.Tensor, band: Band, t0: int = 0, t1: Optional[int] = None, halo: int = 0) -> torch.Tensor:
    lib = _lib.load()
    t1 = band.T if t1 is None else t1
    T_out = t1 - t0
    x = _f32(x)
    assert x.shape[0] == T_out + halo, "x must hold halo + T_out slices"
    if band.b > Band.MAX_KERNEL_BAND:            # wide band: one launch per block of 32 lags, partial sums added
        band._whole_only(t0, t1, halo, "mtransform_dense")
        out = None
        for o, sub in band.chunks():
            part = stencil_fwd(x[: x.shape[0] - o], sub)
            if out is None:
                out = part
            else:
                out[o:].add_(part)
        return out
    NF = x[0].numel() if x.shape[0] else 0
    out = torch.empty((T_out,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    w = band.device_weights(t0, t1, torch.float32)
    _lib.check(lib.tmgcn_mtransform_dense_fwd(_p(x, _F), _p(out, _F), T_out, halo, NF, _p(w, _F), band.b, _stream()))
    return out


def stencil_bwd(g: torch.Tensor, band: Band, t0: int = 0, t1: Optional[int] = None, halo: int = 0) -> torch.Tensor:
    lib = _lib.load()
    t1 = band.T if t1 is None else t1
    T_out = t1 - t0
    g = _f32(g)
    assert g.shape[0] == T_out
    if band.b > Band.MAX_KERNEL_BAND:
        band._whole_only(t0, t1, halo, "mtransform_dense")
        out = None
        for o, sub in band.chunks():
            part = stencil_bwd(g[o:], sub)                 # gradient w.r.t. x[:T-o]
            if out is None:
                out = part
            else:
                out[: g.shape[0] - o].add_(part)
        return out
    NF = g[0].numel() if g.shape[0] else 0
    out = torch.empty((T_out + halo,) + tuple(g.shape[1:]), dtype=torch.float32, device=g.device)
    w = band.device_weights(t0, t1, torch.float32)
    _lib.check(lib.tmgcn_mtransform_dense_bwd(_p(g, _F), _p(out, _F), T_out, halo, NF, _p(w, _F), band.b, _stream()))
    return out


def spmm_raw(csr: SliceCSR, x: torch.Tensor, act: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    x = _f32(x)
    assert x.dim() == 3 and x.shape[0] == csr.T and x.shape[1] == csr.N, "x must be (T, N, F)"
    if csr.val.dtype != torch.float32:
        raise TypeError("spmm: fp32 CSR values only")
    y = torch.empty_like(x) if out is None else out
    _lib.check(lib.tmgcn_spmm_fwd(_p(csr.rowptr, _L), _p(csr.col, _I), _p(csr.val, _F), _p(x, _F), _p(y, _F), csr.T,
                                  csr.N, x.shape[2], act, _stream()))
    return y


def gemm_fwd_raw(p: torch.Tensor, w: torch.Tensor, act: int = 0) -> torch.Tensor:
    lib = _lib.load()
    p, w = _f32(p), _f32(w)
    K, Nf = w.shape
    R = p.numel() // K
    y = torch.empty(tuple(p.shape[:-1]) + (Nf,), dtype=torch.float32, device=p.device)
    _lib.check(lib.tmgcn_gemm_xw_fwd(_p(p, _F), _p(w, _F), _p(y, _F), R, K, Nf, act, _stream()))
    return y


def gemm_bwd_raw(p, w, y, dy, act: int, need_dp: bool = True, need_dw: bool = True):
    lib = _lib.load()
    p, w, dy = _f32(p), _f32(w), _f32(dy)
    y = None if y is None else _f32(y)
    K, Nf = w.shape
    R = dy.numel() // Nf
    dp = torch.empty(tuple(dy.shape[:-1]) + (K,), dtype=torch.float32, device=dy.device) if need_dp else None
    dw = torch.empty_like(w) if need_dw else None
    ws = _ws(lib.tmgcn_gemm_dw_ws_bytes(K, Nf)) if need_dw else None
    _lib.check(lib.tmgcn_gemm_dw_dx_bwd(_p(p), _p(w), _p(y), _p(dy), _p(dp), _p(dw), R, K, Nf, act, _p(ws),
                                        _stream()))
    return dp, dw


class EdgePlan:
    """Flat endpoint ids (ref: ehf:196-198) plus, lazily, the incidence list that
    makes the backward scatter-add deterministic."""

    def __init__(self, edges: torch.Tensor, N: int, t_offset: int = 0, T: Optional[int] = None):
        """edges (3, E): time, src, dst (ref: ehf:169-170); `t_offset` is subtracted from the times (the first
        slice a shard owns); with `T` the times are range-checked too.  The reference raises IndexError on an
        out-of-range edge; here it would be an out-of-bounds device access, so the range is checked once."""
        lib = _lib.load()
        edges = edges.to(device=_dev(), dtype=torch.int64).contiguous()
        if edges.dim() != 2 or edges.shape[0] != 3:
            raise ValueError("edges must be (3, E): time, src, dst")
        self.E = int(edges.shape[1])
        self.n_rows_checked = None
        if self.E:
            lo = edges.min(dim=1).values
            hi = edges.max(dim=1).values
            bad = bool((lo[1:] < 0).any()) or bool((hi[1:] >= N).any()) or int(lo[0]) - t_offset < 0
            if T is not None:
                bad = bad or int(hi[0]) - t_offset >= T
                self.n_rows_checked = T * N
            if bad:
                raise ValueError(f"edge index out of range (N = {N}, T = {T}, first slice = {t_offset})")
        self.src = torch.empty(self.E, dtype=torch.int64, device=edges.device)
        self.dst = torch.empty(self.E, dtype=torch.int64, device=edges.device)
        _lib.check(lib.tmgcn_flat_edge_ids(_p(edges), self.E, N, t_offset, _p(self.src), _p(self.dst), _stream()))
        self._inc = None

    def incidence(self, n_rows: int):
        """(inc_ptr[n_rows+1], perm[2E]): incident (edge, half) codes grouped by endpoint row."""
        if self._inc is None or self._inc[0] != n_rows:
            lib = _lib.load()
            E, dev = self.E, self.src.device
            if E and self.n_rows_checked != n_rows and int(torch.maximum(self.src.max(), self.dst.max())) >= n_rows:
                raise ValueError(f"an edge endpoint lies outside the {n_rows} rows of the embedding tensor")
            keys = torch.cat([self.src, self.dst])
            ar = torch.arange(E, device=dev, dtype=torch.int64)
            code = torch.cat([2 * ar, 2 * ar + 1])
            del ar
            skeys, order = torch.sort(keys, stable=True)
            del keys
            perm = code[order].contiguous()
            del code, order
            inc_ptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
            _lib.check(lib.tmgcn_rowptr_from_sorted_rows(_p(skeys), skeys.numel(), n_rows, _p(inc_ptr), _stream()))
            self._inc = (n_rows, inc_ptr, perm)
        return self._inc[1], self._inc[2]


def readout_fwd_raw(y2d: torch.Tensor, plan: EdgePlan, u: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    y2d, u = _f32(y2d), _f32(u)
    F, Cc = y2d.shape[1], u.shape[1]
    assert u.shape[0] == 2 * F
    out = torch.empty(plan.E, Cc, dtype=torch.float32, device=y2d.device)
    _lib.check(lib.tmgcn_edge_readout_fwd(_p(y2d), _p(plan.src), _p(plan.dst), _p(u), _p(out), plan.E, F, Cc,
                                          _stream()))
    return out


def readout_bwd_raw(y2d, plan: EdgePlan, u, dout, need_dy=True, need_du=True):
    lib = _lib.load()
    y2d, u, dout = _f32(y2d), _f32(u), _f32(dout)
    F, Cc = y2d.shape[1], u.shape[1]
    inc_ptr, perm = plan.incidence(y2d.shape[0])
    dy = torch.empty_like(y2d) if need_dy else None
    du = torch.empty_like(u) if need_du else None
    ws = _ws(lib.tmgcn_edge_readout_bwd_ws_bytes(y2d.shape[0], F, Cc))
    if need_dy or need_du:
        _lib.check(lib.tmgcn_edge_readout_bwd(_p(y2d), _p(u), _p(dout), _p(inc_ptr), _p(perm), _p(dy), _p(du),
                                              y2d.shape[0], F, Cc, 0, _p(ws), _stream()))
    return dy, du


# --------------------------------------------------------------------------
# autograd wiring
# --------------------------------------------------------------------------
class _Stencil(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, band, t0, t1, halo):
        ctx.args = (band, t0, t1, halo)
        return stencil_fwd(_f32(x), band, t0, t1, halo)

    @staticmethod
    def backward(ctx, g):
        band, t0, t1, halo = ctx.args
        return stencil_bwd(_f32(g), band, t0, t1, halo), None, None, None, None


class _Solve(torch.autograd.Function):
    """Y = inv(M) x_3 Z by banded substitution; backward solves with M^T."""

    @staticmethod
    def forward(ctx, z, band):
        lib = _lib.load()
        z = _f32(z)
        y = torch.empty_like(z)
        w = band.device_weights(0, band.T, torch.float32)
        assert z.shape[0] == band.T
        _lib.check(lib.tmgcn_mtransform_dense_solve_fwd(_p(z), _p(y), band.T, z[0].numel(), _p(w), band.b, _stream()))
        ctx.band = band
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        band = ctx.band
        g = _f32(g)
        gz = torch.empty_like(g)
        w = band.device_weights(0, band.T, torch.float32)
        _lib.check(lib.tmgcn_mtransform_dense_solve_bwd(_p(g), _p(gz), band.T, g[0].numel(), _p(w), band.b, _stream()))
        return gz, None


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, csr, act):
        y = spmm_raw(csr, _f32(x), act)
        ctx.csr, ctx.act = csr, act
        if act:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        g = _f32(g)
        if ctx.act:
            (y,) = ctx.saved_tensors
            ge = torch.empty_like(g)
            _lib.check(lib.tmgcn_act_bwd(_p(y), _p(g), _p(ge), g.numel(), ctx.act, _stream()))
            g = ge
        return spmm_raw(ctx.csr.transpose(), g, 0), None, None


class _Gemm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, w, act):
        p, w = _f32(p), _f32(w)
        y = gemm_fwd_raw(p, w, act)
        ctx.act = act
        ctx.save_for_backward(p, w, y if act else None)
        return y

    @staticmethod
    def backward(ctx, g):
        p, w, y = ctx.saved_tensors
        g = _f32(g)
        act = ctx.act
        if act:   # dY <- dY * act'(Y) once, so dP can take the tensor-core path
            lib = _lib.load()
            ge = torch.empty_like(g)
            _lib.check(lib.tmgcn_act_bwd(_p(y), _p(g), _p(ge), g.numel(), act, _stream()))
            g, act = ge, 0
        dp, dw = gemm_bwd_raw(p, w, None, g, act, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dp, dw, None


class _Readout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, u, plan):
        y2d = _f32(y).reshape(-1, y.shape[-1])
        u = _f32(u)
        ctx.plan, ctx.shape = plan, y.shape
        ctx.save_for_backward(y2d, u)
        return readout_fwd_raw(y2d, plan, u)

    @staticmethod
    def backward(ctx, g):
        y2d, u = ctx.saved_tensors
        dy, du = readout_bwd_raw(y2d, ctx.plan, u, _f32(g), ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return (dy.reshape(ctx.shape) if dy is not None else None), du, None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, plan):
        lib = _lib.load()
        F = y.shape[-1]
        y2d = _f32(y).reshape(-1, F)
        z = torch.empty(plan.E, 2 * F, dtype=torch.float32, device=y.device)
        _lib.check(lib.tmgcn_edge_gather_fwd(_p(y2d), _p(plan.src), _p(plan.dst), _p(z), plan.E, F, _stream()))
        ctx.plan, ctx.shape = plan, y.shape
        return z

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        F = ctx.shape[-1]
        n_rows = 1
        for s in ctx.shape[:-1]:
            n_rows *= s
        inc_ptr, perm = ctx.plan.incidence(n_rows)
        dy = torch.empty(n_rows, F, dtype=torch.float32, device=g.device)
        _lib.check(lib.tmgcn_edge_gather_bwd(_p(_f32(g), _F), _p(inc_ptr), _p(perm), _p(dy), n_rows, F, _stream()))
        return dy.reshape(ctx.shape), None


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        lib = _lib.load()
        x = _f32(x)
        y = torch.empty_like(x)
        _lib.check(lib.tmgcn_act_fwd(_p(x), _p(y), x.numel(), act, _stream()))
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        g = _f32(g)
        dx = torch.empty_like(g)
        _lib.check(lib.tmgcn_act_bwd(_p(y), _p(g), _p(dx), g.numel(), ctx.act, _stream()))
        return dx, None


# ---- low-rank backward through the readout factor (see layer_step.py for the algebra) ------------
def class_sums_raw(dout: torch.Tensor, plan: EdgePlan, n_rows: int) -> torch.Tensor:
    """S[row, h, c] = sum over incident (e, h) of dout[e, c]  -> (n_rows, 2C)."""
    lib = _lib.load()
    Cc = dout.shape[1]
    inc_ptr, perm = plan.incidence(n_rows)
    S = torch.empty(n_rows, 2 * Cc, dtype=torch.float32, device=dout.device)
    _lib.check(lib.tmgcn_edge_class_sums(_p(_f32(dout), _F), _p(inc_ptr), _p(perm), _p(S), n_rows, Cc, _stream()))
    return S


def factor_reduce_raw(x2d: torch.Tensor, S: torch.Tensor, Cc: int) -> torch.Tensor:
    """G[(h, f), c] = sum_rows x[row, f] * S[row, h, c]  -> (2F, C)."""
    lib = _lib.load()
    x2d, S = _f32(x2d), _f32(S)
    F = x2d.shape[1]
    G = torch.empty(2 * F, Cc, dtype=torch.float32, device=x2d.device)
    ws = _ws(lib.tmgcn_edge_factor_ws_bytes(F, Cc))
    _lib.check(lib.tmgcn_edge_factor_apply(_p(x2d), _p(G), _p(S), None, _p(G), x2d.shape[0], F, Cc, _p(ws), _stream()))
    return G


def factor_expand_raw(S: torch.Tensor, Vu: torch.Tensor, F: int, Cc: int) -> torch.Tensor:
    """out[row, f] = sum_{h,c} S[row, h, c] * Vu[hF + f, c]  -> (n_rows, F)."""
    lib = _lib.load()
    S, Vu = _f32(S), _f32(Vu)
    out = torch.empty(S.shape[0], F, dtype=torch.float32, device=S.device)
    _lib.check(lib.tmgcn_edge_factor_apply(None, _p(Vu), _p(S), _p(out), None, S.shape[0], F, Cc, None, _stream()))
    return out


def lowrank_small(W: torch.Tensor, U: torch.Tensor, G: torch.Tensor):
    """The (2C x F)-sized algebra of the low-rank backward: U~[(h,c), f] = U[hFo+f, c];
    dW = G~^T U~, dU_h = W^T G_h, V = U~ W^T (returned in the "U layout" factor_expand wants)."""
    Fi, Fo = W.shape
    Cc = U.shape[1]
    J = 2 * Cc
    Ut = U.view(2, Fo, Cc).permute(0, 2, 1).reshape(J, Fo).contiguous()
    Gt = G.view(2, Fi, Cc).permute(0, 2, 1).reshape(J, Fi)
    dW = gemm_fwd_raw(Gt.t().contiguous(), Ut)                                   # (Fi, J) . (J, Fo)
    Wt = W.t().contiguous()
    dU = torch.cat([gemm_fwd_raw(Wt, G[:Fi].contiguous()), gemm_fwd_raw(Wt, G[Fi:].contiguous())])
    V = gemm_fwd_raw(Ut, Wt)                                                     # (J, Fi)
    Vu = V.view(2, Cc, Fi).permute(0, 2, 1).reshape(2 * Fi, Cc).contiguous()
    return dW, dU, Vu


class _PropagateLinearReadout(torch.autograd.Function):
    """out = readout( (A~_t . [M x_3] H) W , U ) for a LINEAR layer (no activation between W and the
    readout: layer 2 of the reference models, ehf:342-355 / 347-355 / 486-495).  Forward = the usual
    kernels; backward = the low-rank chain: nothing F-wide except one pass over P and the final dH."""

    @staticmethod
    def forward(ctx, H, W, U, csr, band, plan):
        H, W, U = _f32(H), _f32(W), _f32(U)
        Ht = stencil_fwd(H, band) if band is not None else H
        P = spmm_raw(csr, Ht) if csr is not None else Ht
        Y = gemm_fwd_raw(P, W)
        out = readout_fwd_raw(Y.reshape(-1, Y.shape[-1]), plan, U)
        ctx.csr, ctx.band, ctx.plan, ctx.shape = csr, band, plan, H.shape
        ctx.save_for_backward(P, W, U)
        return out

    @staticmethod
    def backward(ctx, g):
        P, W, U = ctx.saved_tensors
        g = _f32(g)
        Fi, Cc = W.shape[0], U.shape[1]
        n_rows = P.numel() // Fi
        S = class_sums_raw(g, ctx.plan, n_rows)
        G = factor_reduce_raw(P.reshape(n_rows, Fi), S, Cc)
        dW, dU, Vu = lowrank_small(W, U, G)
        dH = None
        if ctx.needs_input_grad[0]:
            Q = S.view(ctx.shape[0], ctx.shape[1], 2 * Cc)
            if ctx.csr is not None:
                Q = spmm_raw(ctx.csr.transpose(), Q)
            if ctx.band is not None:
                Q = stencil_bwd(Q, ctx.band)
            dH = factor_expand_raw(Q.reshape(n_rows, 2 * Cc), Vu, Fi, Cc).view(ctx.shape)
        return dH, dW, dU, None, None, None


def propagate_linear_readout(H, W, U, csr, band, plan):
    """Differentiable fused linear layer + readout; csr / band may be None (skip SpMM / M-transform)."""
    if U.shape[1] > MAX_FUSED_CLASSES:
        # many classes: the factor kernels hold at most 8 class accumulators -- plain composition of the stages
        Ht = mtransform_dense(H, band) if band is not None else H
        P = spmm(csr, Ht) if csr is not None else Ht
        return edge_readout(gemm_xw(P, W), U, plan)
    return _PropagateLinearReadout.apply(H, W, U, csr, band, plan)


def mtransform_dense(x, band: Band, t0=0, t1=None, halo=0):
    """X~ = X x_3 M (ref: ehf:204), differentiable."""
    return _Stencil.apply(x, band, t0, t1, halo)


def mtransform_dense_inv(z, band: Band):
    """inv(M) x_3 Z (ref: ehf:223-224), differentiable; M banded lower triangular with a nonzero diagonal."""
    if bool((band.w[:, 0] == 0).any()):
        raise ValueError("M has a zero on its diagonal: not invertible")
    if band.b > Band.MAX_KERNEL_BAND:
        raise NotImplementedError(f"inv(M) x_3 Z: the substitution kernel keeps the last b-1 outputs in registers; "
                                  f"b = {band.b} > {Band.MAX_KERNEL_BAND} is not supported")
    return _Solve.apply(z, band)


def spmm(csr: SliceCSR, x, act=None):
    """facewise A[t] @ x[t] (ref: ehf:205-207), differentiable w.r.t. x."""
    return _SpMM.apply(x, csr, ACT[act] if not isinstance(act, int) else act)


def gemm_xw(p, w, act=None):
    """act(p @ w) (ref: ehf:222 + ehf:332-335), differentiable."""
    return _Gemm.apply(p, w, ACT[act] if not isinstance(act, int) else act)


class _GemmSliced(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, w, act):
        lib = _lib.load()
        p, w = _f32(p), _f32(w)
        T, N, K = p.shape
        Nf = w.shape[2]
        y = torch.empty(T, N, Nf, dtype=torch.float32, device=p.device)
        _lib.check(lib.tmgcn_gemm_xw_sliced_fwd(_p(p, _F), _p(w, _F), _p(y, _F), T, N, K, Nf, act, _stream()))
        ctx.act = act
        ctx.save_for_backward(p, w, y if act else None)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        p, w, y = ctx.saved_tensors
        g = _f32(g)
        T, N, K = p.shape
        Nf = w.shape[2]
        dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        _lib.check(lib.tmgcn_gemm_sliced_bwd(_p(p, _F), _p(w, _F), _p(y), _p(g, _F), _p(dp), _p(dw), T, N, K, Nf,
                                             ctx.act, _stream()))
        return dp, dw, None


class _Linear(torch.autograd.Function):
    """y = p @ w + bias in one launch (the bias rides in the GEMM epilogue); backward: dp, dw by the GEMM backward
    kernels and dbias = 1^T dy as a (1 x Nf) dW product."""

    @staticmethod
    def forward(ctx, p, w, bias):
        lib = _lib.load()
        p, w, bias = _f32(p), _f32(w), _f32(bias)
        K, Nf = w.shape
        R = p.numel() // K
        y = torch.empty(tuple(p.shape[:-1]) + (Nf,), dtype=torch.float32, device=p.device)
        _lib.check(lib.tmgcn_gemm_xw_bias_fwd(_p(p, _F), _p(w, _F), _p(bias, _F), _p(y, _F), R, K, Nf, 0, _stream()))
        ctx.save_for_backward(p, w)
        return y

    @staticmethod
    def backward(ctx, g):
        p, w = ctx.saved_tensors
        g = _f32(g)
        dp, dw = gemm_bwd_raw(p, w, None, g, 0, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        db = None
        if ctx.needs_input_grad[2]:
            R = g.numel() // w.shape[1]
            ones = torch.ones(R, 1, dtype=torch.float32, device=g.device)
            _, db = gemm_bwd_raw(ones, torch.empty(1, w.shape[1], device=g.device), None, g, 0, False, True)
            db = db.reshape(-1)
        return dp, dw, db


def linear(p, w, bias):
    """p @ w + bias (ref: nn.Linear at ehf:418-420; w is (K, Nf) = weight^T), differentiable."""
    return _Linear.apply(p, w, bias)


def gemm_xw_sliced(p, w, act=None):
    """per-slice weights: act(p[t] @ w[t]) for w of shape (T, K, Nf) -- the reference's condensed_W=False
    batched matmul (ref: ehf:188-191, 222, 277-282, 330); all slices in one grouped launch (forward, dP and
    dW each), differentiable."""
    assert p.dim() == 3 and w.dim() == 3 and p.shape[0] == w.shape[0] and p.shape[2] == w.shape[1]
    return _GemmSliced.apply(p, w, ACT[act] if not isinstance(act, int) else act)


MAX_FUSED_CLASSES = 8     # the fused readout kernels keep the C class accumulators in registers


def edge_readout(y, u, plan: EdgePlan):
    """[y[src] || y[dst]] @ u (ref: ehf:228-232), differentiable.  Up to 8 classes the classifier is folded into
    the gather (the (E, 2F) concat is never written); wider classifiers take the gather kernel followed by the
    feature GEMM, the reference's own two steps (ehf:228-230, 232)."""
    if u.shape[1] > MAX_FUSED_CLASSES:
        return gemm_xw(edge_gather(y, plan), u)
    return _Readout.apply(y, u, plan)


def edge_gather(y, plan: EdgePlan):
    """[y[src] || y[dst]] (ref: ehf:228-230), differentiable."""
    return _Gather.apply(y, plan)


def activation(x, act):
    return _Act.apply(x, ACT[act] if not isinstance(act, int) else act)
