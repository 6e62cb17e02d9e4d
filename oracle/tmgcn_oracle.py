"""CPU restatement of the TM-GCN propagation hot path (oracle; see package docstring).

Reference = /root/reference/TensorGCN-master (abbreviated ``ref:``);
``ehf`` = ``embedding_help_functions.py``.

Everything here runs on the CPU in the reference's own dtypes: the M-transform
and SpMM in fp64, the slice store / GEMM / gather / classifier in fp32.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = [
    "create_matrix_M",
    "func_MProduct",
    "func_MProduct_dense",
    "split_slices",
    "compute_AX",
    "compute_AtXt",
    "flat_edge_ids",
    "nonlin",
    "apply_Minv",
    "OracleGCN",
    "OracleGCN2",
    "OracleKWGCN",
    "layer_forward",
    "layer_fwd_bwd",
    "normalise_adjacency",
    "compute_f1",
    "create_node_features",
    "split_data",
    "average_precision_class0",
    "row_mrr",
    "compute_MAP_MRR",
    "load_data_mat",
    "make_symmetric",
    "edge_life",
    "laplacian_transformation",
    "create_sparse",
]


# --------------------------------------------------------------------------
# a1. M construction
# --------------------------------------------------------------------------
def create_matrix_M(T: int, no_diag: int, normalize: bool = False) -> torch.Tensor:
    """Banded lower-triangular time-mixing matrix, fp64 (T, T).

    ``normalize=False``: M[t, t-i] = 1/(i+1) for i < no_diag
        (ref: SBM_our.py:88-96, read_data.m:118-124 M_choice=2).
    ``normalize=True``: ones band, each row divided by its sum
        (ref: read_data.py:56-62, full_read_data.py:82-88).
    """
    M = np.zeros((T, T), dtype=np.float64)
    for i in range(min(no_diag, T)):
        w = 1.0 if normalize else 1.0 / (i + 1)
        r = np.arange(i, T)
        M[r, r - i] = w
    if normalize:
        M = M / M.sum(axis=1)[:, None]
    return torch.from_numpy(M)


# --------------------------------------------------------------------------
# a2 / a3. sparse M-transform  A~ = A x_3 M
# --------------------------------------------------------------------------
def func_MProduct(idx, val, shape, M, no_diag=None):
    """Sparse mode-3 product of a coalesced COO tensor with M.

    Follows ref: read_data.py:204-223 (copies SBM_our.py:57-76 ...): for every
    source slice j and every i with M[i, j] != 0 (``nonzero(M[:, j])``,
    :216), slice j's (row, col) pattern is re-stamped with time i and scaled by
    M[i, j]; all pieces are sparse-added and the result is ``coalesce()``d,
    i.e. sorted lexicographically by (t, row, col) with duplicates summed.
    Entries whose sum happens to be 0.0 stay stored (coalesce never prunes).

    idx: (3, nnz) int64, val: (nnz,) float64, shape: (T, N, N), M: (T, T).
    Returns (idx_out (3, nnz_out) int64, val_out (nnz_out,) float64).
    """
    idx = np.asarray(idx, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    T, N, N2 = shape
    assert M.shape[0] == T, "C.size(0) must equal M.size(0)"  # read_data.py:205
    ts, rs, cs, vs = [], [], [], []
    for j in range(T):
        sel = idx[0] == j
        r, c, v = idx[1, sel], idx[2, sel], val[sel]
        tgt = np.nonzero(M[:, j])[0]  # read_data.py:216
        if no_diag is not None:
            assert len(tgt) <= no_diag  # read_data.py:217
        for i in tgt:
            ts.append(np.full(r.shape, i, dtype=np.int64))
            rs.append(r)
            cs.append(c)
            vs.append(M[i, j] * v)
    if not ts:
        return np.zeros((3, 0), np.int64), np.zeros((0,), np.float64)
    t_all = np.concatenate(ts)
    r_all = np.concatenate(rs)
    c_all = np.concatenate(cs)
    v_all = np.concatenate(vs)
    # coalesce(): sort by flattened index, sum equal keys in sorted order
    key = (t_all * N + r_all) * N2 + c_all
    order = np.argsort(key, kind="stable")
    key = key[order]
    v_all = v_all[order]
    first = np.ones(key.shape, dtype=bool)
    first[1:] = key[1:] != key[:-1]
    starts = np.nonzero(first)[0]
    out_val = np.add.reduceat(v_all, starts) if len(starts) else v_all[:0]
    k = key[starts]
    out_idx = np.stack([k // (N * N2), (k // N2) % N, k % N2]).astype(np.int64)
    return out_idx, out_val


def func_MProduct_dense(idx, val, shape, M):
    """Dense route to the same tensor (ref: SBM_our.py:78-86): densify,
    ``M @ B.reshape(T, -1)``, take ``nonzero``.  Small N only.  Differs from
    func_MProduct only where a sum cancels to exactly 0.0 (dropped here)."""
    T, N, N2 = shape
    B = np.zeros(shape, dtype=np.float64)
    np.add.at(B, (idx[0], idx[1], idx[2]), val)
    X = (np.asarray(M, np.float64) @ B.reshape(T, -1)).reshape(shape)
    nz = np.nonzero(X)
    return np.stack(nz).astype(np.int64), X[nz]


# --------------------------------------------------------------------------
# a4. per-slice list
# --------------------------------------------------------------------------
def split_slices(idx, val, T, N=None, dtype=torch.float64):
    """T x N x N COO -> python list of T 2-D sparse COO matrices
    (ref: ehf:561-572, experiment_bitcoin_our.py:53-64).  The reference lets
    torch infer each slice's size from its largest index; passing N pins it
    (identical whenever every slice stores its full diagonal)."""
    idx = torch.as_tensor(np.asarray(idx), dtype=torch.long)
    val = torch.as_tensor(np.asarray(val), dtype=dtype)
    out = []
    for j in range(T):
        sel = idx[0] == j
        if N is None:
            out.append(torch.sparse_coo_tensor(idx[1:3, sel], val[sel]))
        else:
            out.append(torch.sparse_coo_tensor(idx[1:3, sel], val[sel], (N, N)))
    return out


# --------------------------------------------------------------------------
# a5 / a6. dense M-transform + facewise SpMM
# --------------------------------------------------------------------------
def compute_AX(A, X, as_reference: bool = True):
    """AX[k] = A[k] @ X[k] stored into an fp32 (T, N, F) buffer
    (ref: ehf:301-305, ehf:469-473)."""
    T, N = X.shape[0], X.shape[1]
    if as_reference:
        AX = torch.zeros(T, N, X.shape[-1])
        for k in range(len(A)):
            AX[k] = torch.sparse.mm(A[k], X[k])
        return AX
    # same arithmetic without the slice-assign autograd artefact
    return torch.stack([torch.sparse.mm(A[k], X[k]).float() for k in range(len(A))])


def compute_AtXt(At, X, M, as_reference: bool = True):
    """Xt = M @ X.reshape(T, N*F) in fp64, then facewise SpMM rounded to fp32
    at the slice store (ref: ehf:203-208, ehf:307-312)."""
    T = X.shape[0]
    Xt = torch.matmul(M, X.reshape(T, -1)).reshape(X.size())
    return compute_AX(At, Xt, as_reference)


# --------------------------------------------------------------------------
# a8 / a9. nonlinearity and flat edge ids
# --------------------------------------------------------------------------
def nonlin(name: str):
    """ref: ehf:284-289 -- relu / leaky(0.01) / selu."""
    if name == "relu":
        return torch.nn.ReLU()
    if name == "leaky":
        return torch.nn.LeakyReLU(negative_slope=0.01)
    if name == "selu":
        return torch.nn.SELU()
    if name in (None, "none"):
        return torch.nn.Identity()
    raise ValueError(name)


def flat_edge_ids(edges: torch.Tensor, N: int):
    """row id t*N + node for both endpoints (ref: ehf:196-198)."""
    v = torch.tensor([N, 1], dtype=torch.long)
    src = torch.matmul(edges[[0, 1]].transpose(1, 0), v)
    trg = torch.matmul(edges[[0, 2]].transpose(1, 0), v)
    return src, trg


def _readout(Y, src, trg, U):
    """gather both endpoints, concat, classify (ref: ehf:228-232)."""
    Fo = Y.shape[-1]
    Ys = Y.reshape(-1, Fo)[src]
    Yt = Y.reshape(-1, Fo)[trg]
    return torch.matmul(torch.cat((Ys, Yt), dim=1).float(), U)


# --------------------------------------------------------------------------
# a12. module restatements (use_Minv=False, condensed_W=True: the setting of
# every shipped experiment, e.g. experiment_bitcoin_our.py:109)
# --------------------------------------------------------------------------
def apply_Minv(M, Z):
    """Y = inv(M) @ Z.reshape(T, -1) (ref: ehf:183-184, 223-224).  The reference multiplies its fp64 inv(M)
    with the fp32 buffer and raises a dtype error on the shipped fp64 inputs [probed]; this restatement promotes
    to fp64 and rounds the result back to fp32.  Pinned by tests/golden/minv.npz: the UNMODIFIED reference does
    run the flag when M, X and the slices are all fp32 (fp32 LAPACK inverse, fp32 matmul), and this function
    reproduces those outputs and gradients to ~3e-7 (tests/test_oracle_golden.py::test_use_minv_golden).
    EmbeddingGCN2(use_Minv=True) runs in no dtype configuration of the reference (ehf:335 up-casts to fp64 before
    the fp32 inv(M)), so the 2-layer composition of this function stays unpinned."""
    Minv = torch.from_numpy(np.linalg.inv(M.numpy()))
    return torch.matmul(Minv, Z.double().reshape(Z.shape[0], -1)).reshape(Z.size()).float()


class OracleGCN(torch.nn.Module):
    """1-layer TM-GCN (ref: ehf:156-234)."""

    def __init__(self, At, X, edges, M, W, U, as_reference=True, use_Minv=False):
        super().__init__()
        self.use_Minv = use_Minv
        self.M, self.N, self.as_reference = M, X.shape[1], as_reference
        self.W = torch.nn.Parameter(W.clone())
        self.U = torch.nn.Parameter(U.clone())
        self.AtXt = compute_AtXt(At, X, M, as_reference)  # ehf:195
        self.src, self.trg = flat_edge_ids(edges, self.N)

    def forward(self, At=None, X=None, edges=None):
        if type(At) == list and type(X) == torch.Tensor and type(edges) == torch.Tensor:  # ehf:212
            AtXt = compute_AtXt(At, X, self.M, self.as_reference)
            src, trg = flat_edge_ids(edges, self.N)
        else:
            AtXt, src, trg = self.AtXt, self.src, self.trg
        Y = torch.matmul(AtXt, self.W)  # ehf:222
        if self.use_Minv:
            Y = apply_Minv(self.M, Y)  # ehf:223-224
        return _readout(Y, src, trg, self.U)


class OracleGCN2(torch.nn.Module):
    """2-layer TM-GCN (ref: ehf:236-357), use_Minv=False."""

    def __init__(self, At, X, edges, M, W1, W2, U, apply_M_twice=False,
                 apply_M_three_times=False, nonlin2="relu", as_reference=True, use_Minv=False):
        super().__init__()
        self.use_Minv = use_Minv
        self.At, self.M, self.N = At, M, X.shape[1]
        self.apply_M_twice, self.apply_M_three_times = apply_M_twice, apply_M_three_times
        self.as_reference = as_reference
        self.W1 = torch.nn.Parameter(W1.clone())
        self.W2 = torch.nn.Parameter(W2.clone())
        self.U = torch.nn.Parameter(U.clone())
        self.f = nonlin(nonlin2)
        self.AtXt = compute_AtXt(At, X, M, as_reference)
        self.src, self.trg = flat_edge_ids(edges, self.N)

    def forward(self, At=None, X=None, edges=None):
        if type(At) == list and type(X) == torch.Tensor and type(edges) == torch.Tensor:
            AtXt = compute_AtXt(At, X, self.M, self.as_reference)
            src, trg = flat_edge_ids(edges, self.N)
        else:
            AtXt, src, trg = self.AtXt, self.src, self.trg
        if self.use_Minv:  # ehf:331-341
            Y = self.f(apply_Minv(self.M, torch.matmul(AtXt, self.W1))).double()
            AtYt = compute_AtXt(self.At, Y, self.M, self.as_reference)
            Z = apply_Minv(self.M, torch.matmul(AtYt, self.W2))
            return _readout(Z, src, trg, self.U)
        Y = self.f(torch.matmul(AtXt, self.W1)).double()  # ehf:330-335
        if self.apply_M_twice:  # ehf:342-346
            Z = torch.matmul(compute_AtXt(self.At, Y, self.M, self.as_reference), self.W2)
            if self.apply_M_three_times:
                T = Z.shape[0]
                Z = torch.matmul(self.M, Z.reshape(T, -1).double()).reshape(Z.size())
        else:  # ehf:347-349
            Z = torch.matmul(compute_AX(self.At, Y, self.as_reference), self.W2)
        return _readout(Z, src, trg, self.U)  # ehf:351-355


class OracleKWGCN(torch.nn.Module):
    """Static-GCN baseline with 1 or 2 layers (ref: ehf:425-497)."""

    def __init__(self, A, X, edges, W1, U, W2=None, nonlin2="relu", as_reference=True):
        super().__init__()
        self.A, self.N, self.as_reference = A, X.shape[1], as_reference
        self.W1 = torch.nn.Parameter(W1.clone())
        self.W2 = None if W2 is None else torch.nn.Parameter(W2.clone())
        self.U = torch.nn.Parameter(U.clone())
        self.f = nonlin(nonlin2)
        self.src, self.trg = flat_edge_ids(edges, self.N)
        self.AX = compute_AX(A, X, as_reference)

    def forward(self, A=None, X=None, edges=None):
        if type(A) == list and type(X) == torch.Tensor and type(edges) == torch.Tensor:
            AX = compute_AX(A, X, self.as_reference)
            src, trg = flat_edge_ids(edges, self.N)
        else:
            AX, src, trg = self.AX, self.src, self.trg
        if self.W2 is not None:  # ehf:486-487
            Y = self.f(torch.matmul(AX, self.W1)).double()
            Z = torch.matmul(compute_AX(self.A, Y, self.as_reference), self.W2)
        else:
            Z = torch.matmul(AX, self.W1)
        return _readout(Z, src, trg, self.U)


# --------------------------------------------------------------------------
# The benchmarked "layer" (SURVEY.md section 8d): layer 2 of EmbeddingGCN2 with
# apply_M_twice=True (ehf:342-344) + optional nonlinearity + readout and
# classifier (ehf:351-355).
# --------------------------------------------------------------------------
def layer_forward(At, H, M, W, U, edges, act="none", as_reference=True):
    N = H.shape[1]
    P = compute_AtXt(At, H.double(), M, as_reference)  # fp64 math, fp32 store
    Y = nonlin(act)(torch.matmul(P, W))
    src, trg = flat_edge_ids(edges, N)
    return _readout(Y, src, trg, U)


def layer_fwd_bwd(At, H, M, W, U, edges, dOut, act="none", as_reference=True):
    """One layer forward + backward by autograd, exactly how the reference
    obtains its gradients (loss.backward(), experiment_bitcoin_our.py:120).
    Returns out (E, C) fp32 and dH, dW, dU (fp32)."""
    H = H.detach().clone().requires_grad_(True)
    W = W.detach().clone().requires_grad_(True)
    U = U.detach().clone().requires_grad_(True)
    out = layer_forward(At, H, M, W, U, edges, act, as_reference)
    out.backward(dOut)
    return out.detach(), H.grad, W.grad, U.grad


# --------------------------------------------------------------------------
# Input preparation used by the synthetic workloads:
# symmetrise, add I, D^-1/2 (.) D^-1/2 per slice.
# --------------------------------------------------------------------------
def normalise_adjacency(idx, val, T, N):
    """(A + A^T)/2 per slice (ref: read_data.py:98-99), then A + I and
    D^-1/2 (A + I) D^-1/2 with D = row sums (ref: read_data.py:130-164).
    Returns a coalesced (t, i, j)-sorted COO (idx int64, val fp64)."""
    idx = np.asarray(idx, np.int64)
    val = np.asarray(val, np.float64)
    t = np.concatenate([idx[0], idx[0], np.repeat(np.arange(T), N)])
    r = np.concatenate([idx[1], idx[2], np.tile(np.arange(N), T)])
    c = np.concatenate([idx[2], idx[1], np.tile(np.arange(N), T)])
    v = np.concatenate([val * 0.5, val * 0.5, np.ones(T * N)])
    key = (t * N + r) * N + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    first = np.ones(key.shape, bool)
    first[1:] = key[1:] != key[:-1]
    starts = np.nonzero(first)[0]
    v = np.add.reduceat(v, starts)
    key = key[starts]
    t, r, c = key // (N * N), (key // N) % N, key % N
    deg = np.zeros(T * N)
    np.add.at(deg, t * N + r, v)
    dinv = 1.0 / np.sqrt(deg)
    v = v * dinv[t * N + r] * dinv[t * N + c]
    return np.stack([t, r, c]).astype(np.int64), v


# --------------------------------------------------------------------------
# graph preparation (the "next" row in front of the sparse M-transform)
# --------------------------------------------------------------------------
def _coalesce(t, r, c, v, N):
    key = (t * N + r) * N + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    first = np.ones(key.shape, bool)
    first[1:] = key[1:] != key[:-1]
    starts = np.nonzero(first)[0]
    v = np.add.reduceat(v, starts) if len(starts) else v[:0]
    key = key[starts]
    return np.stack([key // (N * N), (key // N) % N, key % N]).astype(np.int64), v


def make_symmetric(idx, val, T, N):
    """(A_t + A_t^T)/2 per slice (ref: read_data.py:88-109)."""
    idx, val = np.asarray(idx, np.int64), np.asarray(val, np.float64)
    return _coalesce(np.concatenate([idx[0], idx[0]]), np.concatenate([idx[1], idx[2]]),
                     np.concatenate([idx[2], idx[1]]), np.concatenate([val, val]) / 2, N)


def edge_life(idx, val, T, N, window):
    """A_new[t] = sum_{s=max(0,t-w+1)}^{t} A[s] (ref: read_data.py:116-125)."""
    idx, val = np.asarray(idx, np.int64), np.asarray(val, np.float64)
    ts, rs, cs, vs = [], [], [], []
    for t in range(T):
        sel = (idx[0] >= max(0, t - window + 1)) & (idx[0] <= t)
        ts.append(np.full(int(sel.sum()), t, np.int64))
        rs.append(idx[1, sel])
        cs.append(idx[2, sel])
        vs.append(val[sel])
    return _coalesce(np.concatenate(ts), np.concatenate(rs), np.concatenate(cs), np.concatenate(vs), N)


def laplacian_transformation(idx, val, T, N):
    """D^-1/2 (B + I) D^-1/2 per slice, D = row sums of B + I (ref: read_data.py:130-164)."""
    idx, val = np.asarray(idx, np.int64), np.asarray(val, np.float64)
    t = np.concatenate([idx[0], np.repeat(np.arange(T), N)])
    r = np.concatenate([idx[1], np.tile(np.arange(N), T)])
    c = np.concatenate([idx[2], np.tile(np.arange(N), T)])
    v = np.concatenate([val, np.ones(T * N)])
    ci, cv = _coalesce(t, r, c, v, N)
    deg = np.zeros(T * N)
    np.add.at(deg, ci[0] * N + ci[1], cv)
    d = 1.0 / np.sqrt(deg)
    return ci, cv * d[ci[0] * N + ci[1]] * d[ci[0] * N + ci[2]]


def create_sparse(idx, val, start, end):
    """time window [start, end), re-based (ref: read_data.py:174-183)."""
    idx, val = np.asarray(idx, np.int64), np.asarray(val, np.float64)
    sel = (idx[0] >= start) & (idx[0] < end)
    out = idx[:, sel].copy()
    out[0] -= start
    return out, val[sel]


# ---------------------------------------------------------------------------------------------------------
# data formats and metrics around the path -- numpy, small inputs only
# ---------------------------------------------------------------------------------------------------------
def compute_f1(guess, target):
    """ref: ehf:530-538 -- class 0 is the positive class."""
    guess, target = np.asarray(guess), np.asarray(target)
    tp = float(np.sum((guess == 0) & (target == 0)))
    fp = float(np.sum((guess == 0) & (target != 0)))
    fn = float(np.sum((guess != 0) & (target == 0)))
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = np.float64(tp) / np.float64(tp + fp)
        recall = np.float64(tp) / np.float64(tp + fn)
        f1 = 2 * (precision * recall) / (precision + recall)
    return precision, recall, f1


def create_node_features(A_idx, A_val, T, N, S_train, S_val, S_test, same_block_size):
    """ref: ehf:597-610 -- column sums and row sums of every slice in an fp32 buffer, blocks returned as fp64."""
    X = np.zeros((T, N, 2), dtype=np.float32)
    for (t, i, j), v in zip(np.asarray(A_idx).T, np.asarray(A_val)):
        X[t, j, 0] += np.float32(v)
        X[t, i, 1] += np.float32(v)
    X = X.astype(np.float64)
    if same_block_size:
        return X[0:S_train], X[S_val:S_train + S_val], X[S_val + S_test:]
    return X[0:S_train], X[S_train:S_train + S_val], X[S_train + S_val:]


def split_data(edges_aug, labels, S_train, S_val, S_test, same_block_size):
    """ref: ehf:612-655 (returns numpy arrays in the reference's order)."""
    edges_aug, labels = np.asarray(edges_aug), np.asarray(labels)

    def cut(lo, hi):
        m = edges_aug[0] >= lo
        if hi is not None:
            m &= edges_aug[0] < hi
        e = edges_aug[:, m].copy()
        e[0] -= lo
        later = e[:, e[0] != 0].copy()
        later[0] -= 1
        return e, labels[m], later
    tr = cut(0, S_train)
    if same_block_size:
        va, te = cut(S_val, S_train + S_val), cut(S_val + S_test, None)
        K_val = int(np.sum(va[0][0] - (S_train - S_val - 1) > 0))
        K_test = int(np.sum(te[0][0] - (S_train - S_test - 1) > 0))
        return (*tr, *va, K_val, *te, K_test)
    va, te = cut(S_train, S_train + S_val), cut(S_train + S_val, None)
    return (*tr, *va, *te)


def average_precision_class0(true_classes, scores):
    """What ehf:711 asks of sklearn: AP = sum_n (R_n - R_{n-1}) P_n over the distinct score thresholds in
    descending order, with class 0 as the positive label."""
    y = (np.asarray(true_classes) == 0)
    s = np.asarray(scores)
    order = np.argsort(-s, kind="mergesort")
    y, s = y[order], s[order]
    n_pos = float(y.sum())
    ap, tp, prev_recall = 0.0, 0.0, 0.0
    k = 0
    while k < len(s):
        e = k
        while e < len(s) and s[e] == s[k]:
            tp += float(y[e])
            e += 1
        recall = tp / n_pos
        ap += (recall - prev_recall) * (tp / e)
        prev_recall = recall
        k = e
    return ap


def row_mrr(pred_row, true_row):
    """ref: ehf:669-681 -- mean reciprocal rank of the label-0 cells of one dense row."""
    existing = np.asarray(true_row) == 0
    order = np.flip(np.argsort(pred_row))
    ranks = np.arange(1, len(pred_row) + 1, dtype=np.float64)[existing[order]]
    return (1.0 / ranks).sum() / ranks.shape[0]


def compute_MAP_MRR(output, target, edges):
    """ref: ehf:684-729 -- per-slice MAP (softmax score of class 0) and MRR (raw column 0 scattered into a dense
    matrix, duplicates added, rows holding a label-1 cell), weighted by the slices' edge counts."""
    output, target, edges = np.asarray(output), np.asarray(target), np.asarray(edges)
    probs = torch.softmax(torch.from_numpy(np.ascontiguousarray(output)), dim=1)[:, 0].numpy()     # ehf:706
    MAP = MRR = 0.0
    for k in np.unique(edges[0]):
        m = edges[0] == k
        w = m.sum() / float(len(m))
        adj = edges[1:3, m]
        pred = np.zeros((adj[0].max() + 1, adj[1].max() + 1), dtype=output.dtype)
        true = np.zeros(pred.shape, dtype=np.int64)
        np.add.at(pred, (adj[0], adj[1]), output[m, 0])
        np.add.at(true, (adj[0], adj[1]), target[m])
        rows = [row_mrr(pred[i], true[i]) for i in range(pred.shape[0]) if (true[i] == 1).any()]
        MRR += (np.mean(rows) if rows else np.nan) * w
        MAP += average_precision_class0(target[m], probs[m]) * w
    return MAP, MRR


def load_data_mat(path, S_train, S_val, S_test, transformed):
    """ref: ehf:542-595 -- the .mat wire format (1-based nnz x 3 subs, nnz x 1 vals) as plain (idx, val) pairs:
    returns (A_labels, blocks, N[, M]) with blocks = three lists of per-slice (idx2, val) pairs."""
    import scipy.io as sio
    saved = sio.loadmat(path)

    def coo(name, shape=None):
        subs = np.asarray(saved[name + "_subs"]).astype(np.int64) - 1
        vals = np.asarray(saved[name + "_vals"], dtype=np.float64).reshape(-1)
        order = np.lexsort((subs[:, 2], subs[:, 1], subs[:, 0]))
        return subs[order].T, vals[order]
    a_idx, a_val = coo("A_labels")
    T, N = int(a_idx[0].max()) + 1, int(max(a_idx[1].max(), a_idx[2].max())) + 1

    def slices(idx, val, lo, hi):
        return [(idx[1:3, idx[0] == j], val[idx[0] == j]) for j in range(lo, hi)]
    if transformed:
        blocks = [slices(*coo(n), 0, S_train) for n in ("Ct_train", "Ct_val", "Ct_test")]
        return (a_idx, a_val), blocks, N, np.asarray(saved["M"], dtype=np.float64)
    c_idx, c_val = coo("C")
    blocks = [slices(c_idx, c_val, 0, S_train), slices(c_idx, c_val, S_train, S_train + S_val),
              slices(c_idx, c_val, S_train + S_val, S_train + S_val + S_test)]
    return (a_idx, a_val), blocks, N
