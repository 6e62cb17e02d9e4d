"""Data formats and metrics on either side of the propagation path.

Same names, argument meaning and return order as the reference's helpers in ``embedding_help_functions.py``
(ehf) so the experiment scripts can switch imports; everything after the ``.mat`` file is read stays on the
device (no per-slice Python masking, no dense N x N score matrices, no host round trips inside the metrics).

    load_data              ehf:542-595   .mat wire format of read_data.m:210-232  ->  device tensors / SliceCSR
    save_mat               read_data.m:210-232 / read_data.py:248-270   the writer of that format
    create_node_features   ehf:597-610   in/out degree features, train / val / test blocks
    split_data             ehf:612-655   edge list + labels -> train / val / test blocks
    compute_f1             ehf:530-538   precision / recall / F1 of class 0
    compute_MAP_MRR        ehf:669-729   per-slice average precision and mean reciprocal rank, weighted
    print_f1               ehf:658-666

These are plumbing around the hot path: sorting, segmented sums and prefix sums are torch device ops; the
CSR-of-slices the models consume is built by the C ABI (``SliceCSR.from_coo``).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .ops import SliceCSR

__all__ = ["load_data", "save_mat", "create_node_features", "split_data", "compute_f1", "compute_MAP_MRR",
           "get_MAP", "get_MRR", "print_f1"]


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


# ---------------------------------------------------------------------------------------------------------
# .mat wire format
# ---------------------------------------------------------------------------------------------------------
def _subs_vals(saved, name: str, one_based: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """``<name>_subs`` (nnz x 3, MATLAB 1-based, any numeric dtype) and ``<name>_vals`` (nnz x 1) -> (3, nnz)
    int64 zero-based indices and a flat fp64 value vector."""
    subs = np.asarray(saved[name + "_subs"])
    vals = np.asarray(saved[name + "_vals"], dtype=np.float64).reshape(-1)
    if subs.ndim != 2 or subs.shape[1] != 3:
        raise ValueError(f"{name}_subs must be nnz x 3, got {subs.shape}")
    if subs.shape[0] != vals.shape[0]:
        raise ValueError(f"{name}_subs has {subs.shape[0]} rows but {name}_vals has {vals.shape[0]}")
    idx = torch.from_numpy(np.ascontiguousarray(subs.T).astype(np.int64)) - (1 if one_based else 0)
    return idx, torch.from_numpy(vals)


def _coalesced(idx: torch.Tensor, val: torch.Tensor, shape, device) -> torch.Tensor:
    return torch.sparse_coo_tensor(idx.to(device), val.to(device), tuple(int(s) for s in shape)).coalesce()


def _csr_window(C: torch.Tensor, lo: int, hi: int, N: int, dtype) -> SliceCSR:
    """slices [lo, hi) of a coalesced (T, N, N) COO tensor as one CSR-of-slices (ehf:561-572 / 581-592 build a
    Python list with one boolean mask pass per slice)."""
    idx, val = C._indices(), C._values()
    t = idx[0]
    # coalesced order is (t, i, j)-lexicographic: the window is one contiguous range
    a = int(torch.searchsorted(t, torch.tensor(lo, device=t.device)))
    b = int(torch.searchsorted(t, torch.tensor(hi, device=t.device)))
    sub = idx[:, a:b].clone()
    sub[0] -= lo
    return SliceCSR.from_coo(sub, val[a:b], hi - lo, N, dtype=dtype)


def load_data(data_loc: str, mat_f_name: str, S_train: int, S_val: int, S_test: int, transformed: bool,
              device=None, dtype=torch.float32, as_list: bool = False):
    """ref: ehf:542-595.  Returns, like the reference,
        transformed=True :  A, A_labels, Ct_train, Ct_val, Ct_test, N, M
        transformed=False:  A, A_labels, C_train,  C_val,  C_test,  N
    A (ones, fp32) and A_labels (fp64) are coalesced sparse COO tensors of shape (T, N, N) on `device`; M is the
    fp64 T x T matrix.  The three adjacency blocks are `SliceCSR` objects (what the models take as `At`; fp32
    values unless `dtype` says otherwise) -- `as_list=True` gives the reference's Python lists of 2-D sparse
    matrices instead.  Like the reference, the transformed blocks all have S_train slices (ehf:563-572) and the
    untransformed C is cut into [0, S_train), [S_train, S_train+S_val), [S_train+S_val, S_train+S_val+S_test).
    """
    import scipy.io as sio
    dev = _device(device)
    saved = sio.loadmat(os.path.join(data_loc, mat_f_name))
    idx, lab = _subs_vals(saved, "A_labels")
    if idx.numel() == 0:
        raise ValueError("A_labels is empty")
    T = int(idx[0].max()) + 1
    N = int(max(idx[1].max(), idx[2].max())) + 1
    A_labels = _coalesced(idx, lab, (T, N, N), dev)
    A = torch.sparse_coo_tensor(A_labels._indices(), torch.ones(A_labels._nnz(), dtype=torch.float32, device=dev),
                                (T, N, N)).coalesce()

    def block(C, lo, hi):
        csr = _csr_window(C, lo, hi, N, dtype)
        if not as_list:
            return csr
        coo_idx, coo_val = csr.to_coo()
        out = []
        for j in range(hi - lo):
            m = coo_idx[0] == j
            out.append(torch.sparse_coo_tensor(coo_idx[1:3, m], coo_val[m], (N, N)).coalesce())
        return out

    if transformed:
        blocks = []
        for name in ("Ct_train", "Ct_val", "Ct_test"):
            ci, cv = _subs_vals(saved, name)
            if ci.numel() and (int(ci[0].max()) >= S_train or int(ci[1:].max()) >= N):
                raise ValueError(f"{name}_subs exceeds the (S_train, N, N) = ({S_train}, {N}, {N}) shape of ehf:551")
            blocks.append(block(_coalesced(ci, cv, (S_train, N, N), dev), 0, S_train))
        M = torch.tensor(np.asarray(saved["M"]), dtype=torch.float64)
        return A, A_labels, blocks[0], blocks[1], blocks[2], N, M
    ci, cv = _subs_vals(saved, "C")
    C = _coalesced(ci, cv, (T, N, N), dev)
    if S_train + S_val + S_test > T:
        raise ValueError(f"S_train+S_val+S_test = {S_train + S_val + S_test} exceeds the {T} stored slices")
    return (A, A_labels, block(C, 0, S_train), block(C, S_train, S_train + S_val),
            block(C, S_train + S_val, S_train + S_val + S_test), N)


def save_mat(path: str, M: Optional[torch.Tensor] = None, **tensors) -> None:
    """Write sparse (T, N, N) tensors in the wire format the experiment scripts load (read_data.m:210-232):
    for every keyword `name=X` (a sparse COO tensor, or a `SliceCSR`) the variables ``name_subs`` (nnz x 3,
    1-based, double -- as MATLAB's sptensor.subs) and ``name_vals`` (nnz x 1, double); plus ``M``."""
    import scipy.io as sio
    out = {}
    for name, X in tensors.items():
        if isinstance(X, SliceCSR):
            idx, val = X.to_coo()
        else:
            X = X if X.is_coalesced() else X.coalesce()
            idx, val = X._indices(), X._values()
        out[name + "_subs"] = idx.t().to(torch.float64).cpu().numpy() + 1.0
        out[name + "_vals"] = val.to(torch.float64).cpu().numpy()[:, None]
    if M is not None:
        out["M"] = torch.as_tensor(M).to(torch.float64).cpu().numpy()
    sio.savemat(path, out, do_compression=True)


# ---------------------------------------------------------------------------------------------------------
# node features and data split
# ---------------------------------------------------------------------------------------------------------
def create_node_features(A: torch.Tensor, S_train: int, S_val: int, S_test: int, same_block_size: bool):
    """ref: ehf:597-610.  X[t, n, 0] = sum_i A[t, i, n] (in-degree), X[t, n, 1] = sum_j A[t, n, j] (out-degree),
    accumulated in fp32 like the reference's `t.zeros` buffer, returned as fp64 blocks:
    same_block_size (the TM-GCN models): train = [0, S_train), val = [S_val, S_train+S_val), test = [S_val+S_test, T);
    otherwise consecutive blocks."""
    A = A if A.is_coalesced() else A.coalesce()
    T, N = A.shape[0], A.shape[1]
    idx, val = A._indices(), A._values().to(torch.float32)
    X = torch.zeros(T * N, 2, dtype=torch.float32, device=val.device)
    X[:, 0].index_add_(0, idx[0] * N + idx[2], val)
    X[:, 1].index_add_(0, idx[0] * N + idx[1], val)
    X = X.view(T, N, 2)
    X_train = X[0:S_train].double()
    if same_block_size:
        return X_train, X[S_val:S_train + S_val].double(), X[S_val + S_test:].double()
    return X_train, X[S_train:S_train + S_val].double(), X[S_train + S_val:].double()


def _block(edges: torch.Tensor, labels: torch.Tensor, lo: int, hi: Optional[int]):
    """edges with time in [lo, hi) re-based to lo, their labels, and the same edges without the block's first
    slice and shifted one slice down (the `e_*` outputs of ehf:618-619)."""
    t = edges[0]
    m = t >= lo
    if hi is not None:
        m = m & (t < hi)
    e = edges[:, m].clone()
    e[0] -= lo
    later = e[:, e[0] != 0].clone()
    later[0] -= 1
    return e, labels[m], later


def split_data(edges_aug: torch.Tensor, labels: torch.Tensor, S_train: int, S_val: int, S_test: int,
               same_block_size: bool):
    """ref: ehf:612-655.  edges_aug (3, E) int64 rows (t, i, j), labels (E,).  Returns
        same_block_size=True : edges_train, target_train, e_train, edges_val, target_val, e_val, K_val,
                               edges_test, target_test, e_test, K_test
        otherwise            : the same without K_val / K_test.
    With equal block sizes the val / test blocks overlap the training block (they start S_val resp.
    S_val+S_test slices in) and K_* counts their edges in the slices the training block has not seen.
    The input is not modified (the reference shifts views of it in place, ehf:627-630)."""
    tr = _block(edges_aug, labels, 0, S_train)
    if same_block_size:
        va = _block(edges_aug, labels, S_val, S_train + S_val)
        te = _block(edges_aug, labels, S_val + S_test, None)
        K_val = torch.sum(va[0][0] - (S_train - S_val - 1) > 0)
        K_test = torch.sum(te[0][0] - (S_train - S_test - 1) > 0)
        return (*tr, *va, K_val, *te, K_test)
    va = _block(edges_aug, labels, S_train, S_train + S_val)
    te = _block(edges_aug, labels, S_train + S_val, None)
    return (*tr, *va, *te)


# ---------------------------------------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------------------------------------
def compute_f1(guess: torch.Tensor, target: torch.Tensor):
    """ref: ehf:530-538.  Class 0 is the positive (minority) class.  fp64 scalars on the inputs' device."""
    g0, t0 = guess == 0, target == 0
    tp = torch.sum(g0 & t0, dtype=torch.float64)
    fp = torch.sum(g0 & ~t0, dtype=torch.float64)
    fn = torch.sum(~g0 & t0, dtype=torch.float64)
    precision = tp / (tp + fp)
    recall = tp / (tp + fn)
    return precision, recall, 2 * (precision * recall) / (precision + recall)


def _segment_starts(keys_sorted: torch.Tensor) -> torch.Tensor:
    """boolean mask of the first element of every run of equal keys (1-D or rows of a 2-D key matrix)."""
    n = keys_sorted.shape[0]
    first = torch.ones(n, dtype=torch.bool, device=keys_sorted.device)
    if n > 1:
        neq = keys_sorted[1:] != keys_sorted[:-1]
        first[1:] = neq if neq.dim() == 1 else neq.any(dim=1)
    return first


def _slice_ids(edges: torch.Tensor):
    """dense ids 0..S-1 of the time slices present in edges[0] (ascending, like `edges[0].unique()`)."""
    uniq, inv = torch.unique(edges[0], sorted=True, return_inverse=True)
    return uniq, inv


def _map_per_slice(score: torch.Tensor, positive: torch.Tensor, sl: torch.Tensor, S: int) -> torch.Tensor:
    """average precision of every slice: sum over the distinct score thresholds n (descending) of
    (R_n - R_{n-1}) * P_n, which is what sklearn's average_precision_score computes (ehf:711)."""
    E = score.numel()
    o1 = torch.argsort(score, descending=True, stable=True)
    o2 = torch.argsort(sl[o1], stable=True)
    order = o1[o2]                                   # by slice, then by descending score
    s_sl, s_sc, s_pos = sl[order], score[order], positive[order].to(torch.float64)
    cnt = torch.bincount(sl, minlength=S)
    start = torch.cumsum(cnt, 0) - cnt
    tps = torch.cumsum(s_pos, 0)
    tps_before = torch.cat([tps.new_zeros(1), tps])[start]          # positives before each slice's first element
    tp = tps - tps_before[s_sl]
    rank = torch.arange(E, device=score.device, dtype=torch.float64) - start[s_sl].to(torch.float64) + 1.0
    # a threshold ends where the (slice, score) pair changes
    last = torch.ones(E, dtype=torch.bool, device=score.device)
    if E > 1:
        last[:-1] = (s_sl[1:] != s_sl[:-1]) | (s_sc[1:] != s_sc[:-1])
    tp_l, rank_l, sl_l = tp[last], rank[last], s_sl[last]
    n_pos = torch.zeros(S, dtype=torch.float64, device=score.device).index_add_(0, s_sl, s_pos)
    first_l = _segment_starts(sl_l)
    prev_tp = torch.where(first_l, torch.zeros_like(tp_l), torch.roll(tp_l, 1))
    contrib = (tp_l - prev_tp) / n_pos[sl_l] * (tp_l / rank_l)
    return torch.zeros(S, dtype=torch.float64, device=score.device).index_add_(0, sl_l, contrib)


def _mrr_per_slice(score: torch.Tensor, label: torch.Tensor, sl: torch.Tensor, row: torch.Tensor, col: torch.Tensor,
                   S: int) -> torch.Tensor:
    """ref: ehf:669-702 without the dense matrices.  Per slice the reference scatters scores and labels into
    dense (max_i+1) x (max_j+1) arrays (duplicates add, absent pairs are 0 / label 0), and for every row holding
    a label-1 cell averages 1/rank over the cells with label 0 -- absent pairs included -- ranked by descending
    score.  Here: the stored cells of a row are ranked by a segmented sort; the W - d absent cells of the row
    (score 0) occupy the ranks right after its positive-score cells, so they contribute a difference of
    harmonic numbers, and stored cells with a negative score are pushed down by W - d."""
    dev = score.device
    f64 = torch.float64
    # 1. coalesce duplicate (slice, i, j) cells: scores add in the score dtype, labels add as integers
    W = torch.zeros(S, dtype=torch.int64, device=dev).scatter_reduce_(0, sl, col + 1, "amax", include_self=True)
    R = torch.zeros(S, dtype=torch.int64, device=dev).scatter_reduce_(0, sl, row + 1, "amax", include_self=True)
    Rmax, Wmax = int(R.max()), int(W.max())
    key = (sl * Rmax + row) * Wmax + col
    ukey, inv = torch.unique(key, sorted=True, return_inverse=True)
    c_score = torch.zeros(ukey.numel(), dtype=score.dtype, device=dev).index_add_(0, inv, score)
    c_label = torch.zeros(ukey.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, label.to(torch.int64))
    c_row = ukey // Wmax                               # (slice, i) id, ascending
    c_sl = c_row // Rmax
    # 2. rank the stored cells inside their row by descending score
    o1 = torch.argsort(c_score, descending=True, stable=True)
    o2 = torch.argsort(c_row[o1], stable=True)
    order = o1[o2]
    r_row, r_score, r_label, r_sl = c_row[order], c_score[order], c_label[order], c_sl[order]
    urow, rinv, d = torch.unique_consecutive(r_row, return_inverse=True, return_counts=True)
    start = torch.cumsum(d, 0) - d
    pos = torch.arange(r_row.numel(), device=dev) - start[rinv] + 1          # 1-based rank among stored cells
    n_rows = urow.numel()
    P = torch.zeros(n_rows, dtype=torch.int64, device=dev).index_add_(0, rinv, (r_score > 0).to(torch.int64))
    Z = W[urow // Rmax] - d                                                    # absent cells of the row
    rank = pos + torch.where(r_score < 0, Z[rinv], torch.zeros_like(pos))
    existing = r_label == 0
    s_stored = torch.zeros(n_rows, dtype=f64, device=dev).index_add_(0, rinv, existing.to(f64) / rank.to(f64))
    n_exist = torch.zeros(n_rows, dtype=torch.int64, device=dev).index_add_(0, rinv, existing.to(torch.int64)) + Z
    H = torch.cat([torch.zeros(1, dtype=f64, device=dev),
                   torch.cumsum(1.0 / torch.arange(1, Wmax + 1, dtype=f64, device=dev), 0)])
    s_absent = H[P + Z] - H[P]
    row_mrr = (s_stored + s_absent) / n_exist.to(f64)
    has_one = torch.zeros(n_rows, dtype=torch.int64, device=dev).index_add_(0, rinv, (r_label == 1).to(torch.int64)) > 0
    # 3. mean over the qualifying rows of each slice (NaN when a slice has none, like the empty mean at ehf:701)
    row_sl = urow // Rmax
    num = torch.zeros(S, dtype=f64, device=dev).index_add_(0, row_sl[has_one], row_mrr[has_one])
    den = torch.zeros(S, dtype=f64, device=dev).index_add_(0, row_sl[has_one], torch.ones_like(row_mrr[has_one]))
    return num / den


def get_MAP(predictions: torch.Tensor, true_classes: torch.Tensor, do_softmax: bool = True) -> torch.Tensor:
    """ref: ehf:704-711 -- average precision of class 0 over one set of edges."""
    probs = torch.softmax(predictions, dim=1)[:, 0] if do_softmax else predictions
    sl = torch.zeros(probs.numel(), dtype=torch.int64, device=probs.device)
    return _map_per_slice(probs, true_classes == 0, sl, 1)[0]


def get_MRR(predictions: torch.Tensor, true_classes: torch.Tensor, adj: torch.Tensor, do_softmax: bool = True):
    """ref: ehf:684-702 -- mean over source nodes of the reciprocal rank of the existing (label 0) pairs."""
    probs = torch.softmax(predictions, dim=1)[:, 0] if do_softmax else predictions[:, 0]
    sl = torch.zeros(probs.numel(), dtype=torch.int64, device=probs.device)
    return _mrr_per_slice(probs, true_classes, sl, adj[0], adj[1], 1)[0]


def compute_MAP_MRR(output: torch.Tensor, target: torch.Tensor, edges: torch.Tensor, do_softmax: bool = True):
    """ref: ehf:714-729.  output (E, C) logits, target (E,) labels (0 = existing edge), edges (3, E) rows
    (t, i, j).  MAP and MRR of every time slice, averaged with weights E_t / E; all slices are evaluated in one
    pass on the device.  As in the reference MAP ranks by softmax(output)[:, 0] and MRR by the raw output[:, 0]
    whatever `do_softmax` says (ehf:725-726).  fp64 scalars."""
    dev = output.device
    edges, target = edges.to(dev), target.to(dev)
    uniq, sl = _slice_ids(edges)
    S = uniq.numel()
    w = torch.bincount(sl, minlength=S).to(torch.float64) / float(sl.numel())
    probs = torch.softmax(output, dim=1)[:, 0]
    ap = _map_per_slice(probs, target == 0, sl, S)
    mrr = _mrr_per_slice(output[:, 0], target, sl, edges[1], edges[2], S)
    return torch.sum(ap * w), torch.sum(mrr * w)


def print_f1(precision_train, recall_train, f1_train, loss_train, precision_val, recall_val, f1_val, loss_val,
             precision_test, recall_test, f1_test, loss_test, alpha=None, tr=None, ep=None, is_final=False):
    """ref: ehf:658-666 (same three lines per call)."""
    head = "FINAL:" if is_final else "alpha/Tr/Ep %.2f/%d/%d." % (alpha, tr, ep)
    rows = (("Train", precision_train, recall_train, f1_train, loss_train),
            ("Val", precision_val, recall_val, f1_val, loss_val),
            ("Test", precision_test, recall_test, f1_test, loss_test))
    for i, (name, p, r, f, l) in enumerate(rows):
        print("%s %s precision/recall/f1 %.16f/%.16f/%.16f. %s loss %.16f.%s"
              % (head, name, p, r, f, name, l, "\n" if i == 2 else ""))
