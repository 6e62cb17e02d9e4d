"""Host-side mirror of the reference's call surface for the hot path.

Same names, argument meaning, dispatch rule and error behaviour as
TensorGCN-master/embedding_help_functions.py (ehf) and the func_MProduct /
create_matrix_M helpers -- backed by libtmgcn_b200.so instead of ATen CPU ops.

Differences a caller can observe (all documented in INTEGRATION.md):
  * results and parameters live on the current CUDA device (fp32);
  * `use_Minv=True` applies inv(M) as a banded substitution in time (M x_3 Y = Z) instead of a dense
    T x T product; `condensed_W=False` (per-slice weights) runs one GEMM launch per slice.  Both
    take the general (dense-gradient) backward.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ops import ACT, Band, EdgePlan, SliceCSR


# --------------------------------------------------------------------------
# L2: M-product operators
# --------------------------------------------------------------------------
def create_matrix_M(T: int, no_diag: int, normalize: bool = False) -> torch.Tensor:
    """Banded lower-triangular M, fp64 (T, T): M[t, t-i] = 1/(i+1), i < no_diag
    (ref: SBM_our.py:88-96); `normalize=True` gives the row-normalised ones band
    of ref: read_data.py:56-62.  Host-side helper (T*b numbers)."""
    M = torch.zeros(T, T, dtype=torch.float64)
    for i in range(min(no_diag, T)):
        M.diagonal(-i).fill_(1.0 if normalize else 1.0 / (i + 1))
    if normalize:
        M = M / M.sum(dim=1, keepdim=True)
    return M


def func_MProduct(C: torch.Tensor, M: torch.Tensor, no_diag: Optional[int] = None) -> torch.Tensor:
    """Sparse mode-3 product C x_3 M (ref: read_data.py:204-223).

    C: sparse COO (T, N, N); M: (T, T) banded lower triangular.  Returns a
    coalesced sparse COO tensor (int64 indices in (t, i, j) order, fp64 values)
    on the CUDA device -- index-for-index what the reference's
    `C_new.coalesce()` holds."""
    assert C.size()[0] == M.size()[0]  # ref: read_data.py:205
    T, N, N2 = C.shape
    if N != N2:
        raise ValueError("C must be T x N x N")
    band = Band(M)
    if no_diag is not None:
        assert band.b <= no_diag  # ref: read_data.py:217
    Cc = C if C.is_coalesced() else C.coalesce()
    csr = SliceCSR.from_coo(Cc._indices(), Cc._values(), T, N, dtype=torch.float64)
    out = ops.mtransform_sparse(csr, band)
    idx, val = out.to_coo()
    return torch.sparse_coo_tensor(idx, val, (T, N, N), is_coalesced=True)


def split_slices(Ct: torch.Tensor) -> List[torch.Tensor]:
    """T x N x N sparse COO -> list of T 2-D sparse COO matrices, the `At`
    argument of the modules (ref: ehf:561-572, experiment_bitcoin_our.py:53-56)."""
    Ct = Ct if Ct.is_coalesced() else Ct.coalesce()
    T, N = Ct.shape[0], Ct.shape[1]
    idx, val = Ct._indices(), Ct._values()
    bounds = torch.searchsorted(idx[0].contiguous(), torch.arange(T + 1, device=idx.device))
    out = []
    for j in range(T):
        a, b = int(bounds[j]), int(bounds[j + 1])
        out.append(torch.sparse_coo_tensor(idx[1:3, a:b], val[a:b], (N, N), is_coalesced=True))
    return out


# --------------------------------------------------------------------------
# L3: model modules
# --------------------------------------------------------------------------
def _act_name(nonlin2: str) -> str:
    if nonlin2 not in ("relu", "leaky", "selu"):
        raise ValueError('nonlin2 must be "relu", "leaky" or "selu"')  # ref: ehf:284-289
    return nonlin2


class _Base(nn.Module):
    def _csr(self, A, N) -> SliceCSR:
        """list-of-COO -> device CSR, cached by list identity."""
        if isinstance(A, SliceCSR):
            return A
        cache = self.__dict__.setdefault("_csr_cache", {})
        key = id(A)
        hit = cache.get(key)
        # a hit needs the same list object holding the same slice objects (a list mutated in place is rebuilt)
        if hit is not None and hit[0] is A and len(hit[1]) == len(A) and all(x is y for x, y in zip(hit[1], A)):
            return hit[2]
        csr = SliceCSR.from_slice_list(A, N)
        while len(cache) >= 4:                      # train / val / test lists + one spare: bounded, oldest first
            cache.pop(next(iter(cache)))
        cache[key] = (A, list(A), csr)
        return csr

    @staticmethod
    def _x(X) -> torch.Tensor:
        return X.detach().to(device=ops._dev(), dtype=torch.float32).contiguous()

    @staticmethod
    def _fresh(At, X, edges) -> bool:
        # the reference's dispatch rule (ref: ehf:212, 316, 476): a Python list of slices means "new inputs";
        # a ready SliceCSR (what data.load_data returns) counts as one too
        return ((type(At) == list or isinstance(At, SliceCSR)) and type(X) == torch.Tensor
                and type(edges) == torch.Tensor)

    @staticmethod
    def _param(*shape) -> nn.Parameter:
        # drawn on the CPU generator in the reference's order, so a seed reproduces its init
        return nn.Parameter(torch.randn(*shape).to(ops._dev()))


class EmbeddingGCN(_Base):
    """1-layer TM-GCN (ref: ehf:156-234)."""

    def __init__(self, At, X, edges, M, hidden_feat=[2, 2], condensed_W=False, use_Minv=True):
        super().__init__()
        self.M = M
        self.band = Band(M)
        self.use_Minv = use_Minv
        self.condensed_W = condensed_W
        self.T = X.shape[0]
        self.N = X.shape[1]
        self.F = [X.shape[-1]] + hidden_feat
        if condensed_W:                                       # ref: ehf:188-191
            self.W = self._param(self.F[0], self.F[1])
        else:
            self.W = self._param(self.T, self.F[0], self.F[1])
        self.U = self._param(2 * self.F[1], self.F[2])
        self.AtXt = self.compute_AtXt(At, X)                  # ref: ehf:195
        self.edge_plan = EdgePlan(edges, self.N, T=self.T)              # ref: ehf:196-198

    def compute_AtXt(self, At, X):
        """(T, N, F) fp32 = facewise At[k] @ (M x_3 X)[k] (ref: ehf:203-208)."""
        csr = self._csr(At, self.N)
        with torch.no_grad():
            return ops.spmm_raw(csr, ops.stencil_fwd(self._x(X), self.band))

    def forward(self, At=None, X=None, edges=None):
        if self._fresh(At, X, edges):
            AtXt = self.compute_AtXt(At, X)
            plan = EdgePlan(edges, self.N, T=int(X.shape[0]))
        else:
            AtXt, plan = self.AtXt, self.edge_plan
        if self.use_Minv or not self.condensed_W:             # general path (ehf:222-232)
            Y = ops.gemm_xw(AtXt, self.W) if self.condensed_W else ops.gemm_xw_sliced(AtXt, self.W)
            if self.use_Minv:                                 # Y = inv(M) x_3 (AtXt W), ehf:223-224
                Y = ops.mtransform_dense_inv(Y, self.band)
            return ops.edge_readout(Y, self.U, plan)
        # ref: ehf:222 (GEMM) + ehf:228-232 (readout): a linear map followed by a C-class readout
        return ops.propagate_linear_readout(AtXt, self.W, self.U, None, None, plan)


class EmbeddingGCN2(_Base):
    """2-layer TM-GCN (ref: ehf:236-357)."""

    def __init__(self, At, X, edges, M, hidden_feat=[2, 2, 2], condensed_W=False, use_Minv=True,
                 apply_M_twice=False, apply_M_three_times=False, nonlin2="relu"):
        super().__init__()
        self.condensed_W = condensed_W
        self.At = At
        self.M = M
        self.band = Band(M)
        self.use_Minv = use_Minv
        self.apply_M_twice = apply_M_twice
        self.apply_M_three_times = apply_M_three_times
        self.T = X.shape[0]
        self.N = X.shape[1]
        self.F = [X.shape[-1]] + hidden_feat
        if condensed_W:                                       # ref: ehf:277-282
            self.W1 = self._param(self.F[0], self.F[1])
            self.W2 = self._param(self.F[1], self.F[2])
        else:
            self.W1 = self._param(self.T, self.F[0], self.F[1])
            self.W2 = self._param(self.T, self.F[1], self.F[2])
        self.U = self._param(self.F[2] * 2, self.F[3])
        self.nonlin2 = _act_name(nonlin2)
        self.At_csr = self._csr(At, self.N)
        self.AtXt = self.compute_AtXt(At, X)
        self.edge_plan = EdgePlan(edges, self.N, T=self.T)

    def compute_AX(self, A, X):
        """facewise A[k] @ X[k] (ref: ehf:301-305); differentiable w.r.t. X."""
        return ops.spmm(self._csr(A, self.N), X)

    def compute_AtXt(self, At, X):
        """facewise At[k] @ (M x_3 X)[k] (ref: ehf:307-312); differentiable w.r.t. X
        when X is a CUDA tensor that requires grad."""
        csr = self._csr(At, self.N)
        if X.is_cuda and X.requires_grad:
            return ops.spmm(csr, ops.mtransform_dense(X, self.band))
        with torch.no_grad():
            return ops.spmm_raw(csr, ops.stencil_fwd(self._x(X), self.band))

    def forward(self, At=None, X=None, edges=None):
        if self._fresh(At, X, edges):
            AtXt = self.compute_AtXt(At, X)
            plan = EdgePlan(edges, self.N, T=int(X.shape[0]))
        else:
            AtXt, plan = self.AtXt, self.edge_plan
        if self.use_Minv or not self.condensed_W:             # general path, ehf:330-349 branch by branch
            mm = ops.gemm_xw if self.condensed_W else ops.gemm_xw_sliced
            if self.use_Minv:                                 # ehf:331-332, 337-341
                Y = ops.activation(ops.mtransform_dense_inv(mm(AtXt, self.W1), self.band), self.nonlin2)
                AtYt = ops.spmm(self.At_csr, ops.mtransform_dense(Y, self.band))
                Z = ops.mtransform_dense_inv(mm(AtYt, self.W2), self.band)
                return ops.edge_readout(Z, self.U, plan)
            Y = mm(AtXt, self.W1, self.nonlin2)
            Yt = ops.mtransform_dense(Y, self.band) if self.apply_M_twice else Y
            Z = mm(ops.spmm(self.At_csr, Yt), self.W2)
            if self.apply_M_twice and self.apply_M_three_times:
                Z = ops.mtransform_dense(Z, self.band)
            return ops.edge_readout(Z, self.U, plan)
        Y = ops.gemm_xw(AtXt, self.W1, self.nonlin2)          # layer 1 (ref: ehf:330-335)
        if self.apply_M_twice and self.apply_M_three_times:   # ref: ehf:342-346
            Z = ops.gemm_xw(ops.spmm(self.At_csr, ops.mtransform_dense(Y, self.band)), self.W2)
            Z = ops.mtransform_dense(Z, self.band)
            return ops.edge_readout(Z, self.U, plan)          # ref: ehf:351-355
        # layer 2 is linear up to the C-class readout (ref: ehf:342-344 / 347-349, 351-355): fused op whose
        # backward runs on the rank-2C factor of the readout gradient
        return ops.propagate_linear_readout(Y, self.W2, self.U, self.At_csr,
                                            self.band if self.apply_M_twice else None, plan)


class EmbeddingGCN_reg(_Base):
    """1-layer TM-GCN with a node-regression head (ref: ehf:359-423): out[t, n] = lin1(AtXt[t, n] @ W).
    As in the reference, forward() ignores its arguments and always uses the constructor's inputs."""

    def __init__(self, At, X, M, hidden_feat=[2, 2], condensed_W=False, use_Minv=True):
        super().__init__()
        self.M = M
        self.band = Band(M)
        self.use_Minv = use_Minv
        self.condensed_W = condensed_W
        self.T = X.shape[0]
        self.N = X.shape[1]
        self.F = [X.shape[-1]] + hidden_feat
        if condensed_W:
            self.W = self._param(self.F[0], self.F[1])
        else:
            self.W = self._param(self.T, self.F[0], self.F[1])
        self.lin1 = nn.Linear(self.F[1], 1).to(ops._dev())    # same init draws as the reference's nn.Linear
        csr = self._csr(At, self.N)
        with torch.no_grad():
            self.AtXt = ops.spmm_raw(csr, ops.stencil_fwd(self._x(X), self.band))

    def forward(self, At=None, X=None):
        Y = ops.gemm_xw(self.AtXt, self.W) if self.condensed_W else ops.gemm_xw_sliced(self.AtXt, self.W)
        if self.use_Minv:                                                            # ehf:415-417
            Y = ops.mtransform_dense_inv(Y, self.band)
        out = ops.linear(Y, self.lin1.weight.t().contiguous(), self.lin1.bias)       # (T, N, 1), ehf:418-420
        return out.squeeze(2)


class EmbeddingKWGCN(_Base):
    """Static-GCN baseline, 1 or 2 layers (ref: ehf:425-497)."""

    def __init__(self, A, X, edges, hidden_feat=[2, 2], nonlin2="relu"):
        super().__init__()
        self.no_layers = len(hidden_feat) - 1
        self.T = len(A)
        self.N = X.shape[1]
        self.A = A
        self.F = [X.shape[-1]] + hidden_feat
        if self.no_layers == 2:                               # reference draws W2 first (ehf:451-453)
            self.W2 = self._param(self.F[1], self.F[2])
        self.W1 = self._param(self.F[0], self.F[1])
        self.U = self._param(self.F[-2] * 2, self.F[-1])
        self.nonlin2 = _act_name(nonlin2)
        self.A_csr = self._csr(A, self.N)
        self.edge_plan = EdgePlan(edges, self.N, T=self.T)
        self.AX = self.compute_AX(A, X)

    def compute_AX(self, A, X):
        csr = self._csr(A, self.N)
        if X.is_cuda and X.requires_grad:
            return ops.spmm(csr, X)
        with torch.no_grad():
            return ops.spmm_raw(csr, self._x(X))

    def forward(self, A=None, X=None, edges=None):
        if self._fresh(A, X, edges):
            AX = self.compute_AX(A, X)
            plan = EdgePlan(edges, self.N, T=int(X.shape[0]))
        else:
            AX, plan = self.AX, self.edge_plan
        if self.no_layers == 2:                               # ref: ehf:486-487, 491-495
            Y = ops.gemm_xw(AX, self.W1, self.nonlin2)
            return ops.propagate_linear_readout(Y, self.W2, self.U, self.A_csr, None, plan)
        return ops.propagate_linear_readout(AX, self.W1, self.U, None, None, plan)


class TMGCNLayer(nn.Module):
    """One TM-GCN propagation layer with readout -- the unit the benchmark times
    (SURVEY.md section 8d): H -> M x_3 H -> A~_t . H~_t -> act(. W) -> edge readout . U.
    Equals layer 2 of EmbeddingGCN2(apply_M_twice=True) plus readout/classifier
    (ref: ehf:342-344, 351-355)."""

    def __init__(self, At: SliceCSR, band: Band, edge_plan: EdgePlan, W: torch.Tensor, U: torch.Tensor, act=None,
                 t0: int = 0, t1: Optional[int] = None, halo: int = 0):
        super().__init__()
        self.At, self.band, self.edge_plan, self.act = At, band, edge_plan, act
        self.t0, self.t1, self.halo = t0, (band.T if t1 is None else t1), halo
        self.W = nn.Parameter(W.detach().to(ops._dev(), torch.float32).contiguous())
        self.U = nn.Parameter(U.detach().to(ops._dev(), torch.float32).contiguous())

    def forward(self, H: torch.Tensor, fused: bool = False) -> torch.Tensor:
        """fused=True (linear layer, unsharded only) takes the low-rank-backward op."""
        if fused and not self.act and self.halo == 0 and self.t0 == 0 and self.t1 == self.band.T:
            return ops.propagate_linear_readout(H, self.W, self.U, self.At, self.band, self.edge_plan)
        Ht = ops.mtransform_dense(H, self.band, self.t0, self.t1, self.halo)
        Y = ops.gemm_xw(ops.spmm(self.At, Ht), self.W, self.act)
        return ops.edge_readout(Y, self.U, self.edge_plan)
