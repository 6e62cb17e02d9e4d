#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no
reference tree, it only reads the .npz files this script wrote):

    python tests/golden/make_golden.py

* ``embedding_help_functions`` is imported as-is (behind a stub for the unused
  ``matplotlib`` import, ehf:11).
* ``func_MProduct`` / ``func_MProduct_dense`` / ``create_matrix_M`` live in
  scripts that execute at import (SBM_our.py needs ``dynamicgem``), so their
  source lines are read from /root/reference at run time and exec()'d -- the
  text is never stored in this repo.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch as t

warnings.filterwarnings("ignore")
REF = "/root/reference/TensorGCN-master"
OUT = os.path.dirname(os.path.abspath(__file__))

for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, REF)
import embedding_help_functions as ehf  # noqa: E402


def load_ref_functions(no_diag):
    src = open(os.path.join(REF, "SBM_our.py")).read().split("\n")
    a = [i for i, l in enumerate(src) if l.startswith("def func_MProduct(")][0]
    b = [i for i, l in enumerate(src) if l.startswith("def load_data(")][0]
    ns = {"t": t, "np": np, "no_diag": no_diag}
    exec("\n".join(src[a:b]), ns)
    return ns


def ref_M_normalised(T, no_diag):
    """read_data.py:56-62 executed from the reference text."""
    src = open(os.path.join(REF, "read_data.py")).read().split("\n")
    a = [i for i, l in enumerate(src) if l.startswith("#Create M")][0]
    ns = {"np": np, "torch": t, "T": T, "no_diag": no_diag}
    exec("\n".join(src[a + 1:a + 8]), ns)
    return ns["M"]


def random_coo(T, N, density, seed, symmetric_diag=True):
    g = t.Generator().manual_seed(seed)
    dense = (t.rand(T, N, N, generator=g) < density).double() * (t.rand(T, N, N, generator=g).double() + 0.1)
    if symmetric_diag:
        dense = (dense + dense.transpose(1, 2)) / 2
        dense = dense + t.eye(N, dtype=t.float64)[None]
        d = dense.sum(2)
        dense = dense / t.sqrt(d)[:, :, None] / t.sqrt(d)[:, None, :]
    return dense.to_sparse().coalesce()


def slice_list(C, T):
    # experiment_bitcoin_our.py:53-56 (legacy ctor keeps the fp64 values)
    out = []
    for j in range(T):
        idx = C._indices()[0] == j
        out.append(t.sparse.FloatTensor(C._indices()[1:3, idx], C._values()[idx]))
    return out


def gen_mproduct():
    cases = {}
    for name, (T, N, dens, b, norm, seed) in {
        "t8n50b3": (8, 50, 0.05, 3, False, 1),
        "t8n50b3norm": (8, 50, 0.05, 3, True, 2),
        "t12n33b20": (12, 33, 0.10, 20, False, 3),   # band wider than T
        "t5n17b1": (5, 17, 0.20, 1, False, 4),       # M = I
        "t9n40raw": (9, 40, 0.08, 4, False, 5),      # unsymmetric, no diagonal
    }.items():
        ns = load_ref_functions(b)
        C = random_coo(T, N, dens, seed, symmetric_diag=(name != "t9n40raw"))
        M = ref_M_normalised(T, b) if norm else ns["create_matrix_M"](T, b)
        out = ns["func_MProduct"](C, M)
        outd = ns["func_MProduct_dense"](C, M).coalesce()
        assert t.equal(out._indices(), outd._indices())
        cases[name + "_in_idx"] = C._indices().numpy()
        cases[name + "_in_val"] = C._values().numpy()
        cases[name + "_shape"] = np.array([T, N, N])
        cases[name + "_b"] = np.array(b)
        cases[name + "_M"] = M.numpy()
        cases[name + "_out_idx"] = out._indices().numpy()
        cases[name + "_out_val"] = out._values().numpy()
        cases[name + "_outd_val"] = outd._values().numpy()
    np.savez_compressed(os.path.join(OUT, "mproduct.npz"), **cases)
    print("mproduct.npz", len(cases))


def gen_chess():
    """Real-shape input: first 6 monthly slices of the shipped KONECT chess
    file (whitespace separated, `%` comments), restricted to the 600 busiest
    players so the fixture stays small; symmetrised + normalised the reference
    way, then func_MProduct with b=3."""
    raw = np.loadtxt(os.path.join(REF, "data/chess/out.chess.csv"), comments="%")
    dates = np.unique(raw[:, 3])
    keep = np.isin(raw[:, 3], dates[:40])
    raw = raw[keep]
    # the 6 busiest months of the first 40
    cnt = np.array([(raw[:, 3] == d).sum() for d in dates[:40]])
    months = np.sort(np.argsort(-cnt)[:6])
    ids, c = np.unique(np.concatenate([raw[:, 0], raw[:, 1]]), return_counts=True)
    top = ids[np.argsort(-c)[:600]]
    remap = {int(v): i for i, v in enumerate(np.sort(top))}
    T, N = 6, 600
    dense = t.zeros(T, N, N, dtype=t.float64)
    for k, mth in enumerate(months):
        rows = raw[raw[:, 3] == dates[mth]]
        for a, b_, w, _ in rows:
            if int(a) in remap and int(b_) in remap and a != b_:
                dense[k, remap[int(a)], remap[int(b_)]] = 1.0
    dense = (dense + dense.transpose(1, 2)) / 2 + t.eye(N, dtype=t.float64)[None]
    d = dense.sum(2)
    dense = dense / t.sqrt(d)[:, :, None] / t.sqrt(d)[:, None, :]
    C = dense.to_sparse().coalesce()
    ns = load_ref_functions(3)
    M = ns["create_matrix_M"](T, 3)
    out = ns["func_MProduct"](C, M)
    np.savez_compressed(
        os.path.join(OUT, "chess6.npz"),
        in_idx=C._indices().numpy().astype(np.int32), in_val=C._values().numpy(),
        shape=np.array([T, N, N]), b=np.array(3), M=M.numpy(),
        out_idx=out._indices().numpy().astype(np.int32), out_val=out._values().numpy())
    print("chess6.npz nnz in/out", C._nnz(), out._nnz())


def grads(mod, out, dOut, names):
    for p in mod.parameters():
        p.grad = None
    out.backward(dOut)
    return {n: getattr(mod, n).grad.detach().numpy().copy() for n in names}


def gen_models():
    ns = load_ref_functions(3)
    T, N, F0, E = 6, 40, 2, 90
    C = random_coo(T, N, 0.08, 11)
    M = ns["create_matrix_M"](T, 3)
    Ct = ns["func_MProduct"](C, M)
    At = slice_list(Ct, T)
    A = slice_list(C, T)
    g = t.Generator().manual_seed(12)
    X = t.rand(T, N, F0, generator=g, dtype=t.float64)
    X2 = t.rand(T, N, F0, generator=g, dtype=t.float64)
    edges = t.stack([t.randint(0, T, (E,), generator=g), t.randint(0, N, (E,), generator=g),
                     t.randint(0, N, (E,), generator=g)])
    edges = edges[:, t.argsort(edges[0], stable=True)]
    edges2 = edges[:, t.randperm(E, generator=g)[:50]]
    d = {"C_idx": C._indices().numpy(), "C_val": C._values().numpy(),
         "Ct_idx": Ct._indices().numpy(), "Ct_val": Ct._values().numpy(),
         "M": M.numpy(), "X": X.numpy(), "X2": X2.numpy(),
         "edges": edges.numpy(), "edges2": edges2.numpy(), "TN": np.array([T, N])}

    # 1-layer (ehf:156-234) -- cached and fresh-input forward, gradients
    t.manual_seed(100)
    m = ehf.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    dOut = t.randn(E, 2, generator=g)
    out = m()
    d["gcn1_W"], d["gcn1_U"] = m.W.detach().numpy().copy(), m.U.detach().numpy().copy()
    d["gcn1_AtXt"] = m.AtXt.numpy()
    d["gcn1_out"], d["gcn1_dOut"] = out.detach().numpy(), dOut.numpy()
    for k, v in grads(m, out, dOut, ["W", "U"]).items():
        d["gcn1_d" + k] = v
    with t.no_grad():
        d["gcn1_out_fresh"] = m(At, X2, edges2).numpy()

    # 2-layer variants (ehf:236-357)
    for tag, kw in {
        "relu": dict(nonlin2="relu"),
        "leaky": dict(nonlin2="leaky"),
        "selu": dict(nonlin2="selu"),
        "selu_m2": dict(nonlin2="selu", apply_M_twice=True),
        "relu_m3": dict(nonlin2="relu", apply_M_twice=True, apply_M_three_times=True),
    }.items():
        t.manual_seed(200)
        m = ehf.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=True, use_Minv=False, **kw)
        out = m()
        p = "gcn2_" + tag + "_"
        d[p + "W1"], d[p + "W2"], d[p + "U"] = (x.detach().numpy().copy() for x in (m.W1, m.W2, m.U))
        d[p + "out"] = out.detach().numpy()
        for k, v in grads(m, out, dOut, ["W1", "W2", "U"]).items():
            d[p + "d" + k] = v
        with t.no_grad():
            d[p + "out_fresh"] = m(At, X2, edges2).numpy()

    # static baseline (ehf:425-497), 1 and 2 layers
    for tag, hf in {"kw1": [6, 2], "kw2": [6, 6, 2]}.items():
        t.manual_seed(300)
        m = ehf.EmbeddingKWGCN(A, X, edges, hidden_feat=hf, nonlin2="leaky")
        out = m()
        p = tag + "_"
        names = ["W1", "U"] + (["W2"] if len(hf) == 3 else [])
        for n in names:
            d[p + n] = getattr(m, n).detach().numpy().copy()
        d[p + "out"] = out.detach().numpy()
        for k, v in grads(m, out, dOut, names).items():
            d[p + "d" + k] = v
        with t.no_grad():
            d[p + "out_fresh"] = m(A, X2, edges2).numpy()

    # wide-feature layer: F 32 -> 48 -> 16 -> 3, gradient w.r.t. the layer-2 input
    # is exercised through W1 (chain through compute_AtXt with grad, ehf:343)
    t.manual_seed(400)
    Xw = t.rand(T, N, 32, generator=g, dtype=t.float64)
    m = ehf.EmbeddingGCN2(At, Xw, edges, M, hidden_feat=[48, 16, 3], condensed_W=True, use_Minv=False,
                          apply_M_twice=True, nonlin2="relu")
    # randn weights at F=32..48 blow the activations up; scale them like the tests will
    with t.no_grad():
        m.W1 *= 0.2
        m.W2 *= 0.2
    dOut3 = t.randn(E, 3, generator=g)
    out = m()
    d["wide_X"] = Xw.numpy()
    d["wide_W1"], d["wide_W2"], d["wide_U"] = (x.detach().numpy().copy() for x in (m.W1, m.W2, m.U))
    d["wide_out"], d["wide_dOut"] = out.detach().numpy(), dOut3.numpy()
    for k, v in grads(m, out, dOut3, ["W1", "W2", "U"]).items():
        d["wide_d" + k] = v
    # per-slice weights (condensed_W=False, ehf:188-191 / 277-282) and the regression head (ehf:359-423)
    t.manual_seed(500)
    m = ehf.EmbeddingGCN(At, X, edges, M, hidden_feat=[6, 2], condensed_W=False, use_Minv=False)
    out = m()
    d["gcn1u_W"], d["gcn1u_U"], d["gcn1u_out"] = m.W.detach().numpy().copy(), m.U.detach().numpy().copy(), out.detach().numpy()
    for k, v in grads(m, out, dOut, ["W", "U"]).items():
        d["gcn1u_d" + k] = v
    t.manual_seed(600)
    m = ehf.EmbeddingGCN2(At, X, edges, M, hidden_feat=[6, 6, 2], condensed_W=False, use_Minv=False,
                          apply_M_twice=True, nonlin2="leaky")
    out = m()
    for n in ("W1", "W2", "U"):
        d["gcn2u_" + n] = getattr(m, n).detach().numpy().copy()
    d["gcn2u_out"] = out.detach().numpy()
    for k, v in grads(m, out, dOut, ["W1", "W2", "U"]).items():
        d["gcn2u_d" + k] = v
    t.manual_seed(700)
    m = ehf.EmbeddingGCN_reg(At, X, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    out = m()
    dReg = t.randn(out.shape, generator=g)
    d["reg_W"], d["reg_lw"], d["reg_lb"] = (m.W.detach().numpy().copy(), m.lin1.weight.detach().numpy().copy(),
                                            m.lin1.bias.detach().numpy().copy())
    d["reg_out"], d["reg_dOut"] = out.detach().numpy(), dReg.numpy()
    for p_ in m.parameters():
        p_.grad = None
    out.backward(dReg)
    d["reg_dW"], d["reg_dlw"], d["reg_dlb"] = m.W.grad.numpy().copy(), m.lin1.weight.grad.numpy().copy(), m.lin1.bias.grad.numpy().copy()
    # use_Minv=True (ehf:183-184, 223-224): the reference raises "expected m1 and m2 to have the same dtype, but
    # got: double != float" at ehf:224 (fp64 inv(M) times the fp32 AtXt buffer) on these fp64 inputs; the flag is
    # pinned by gen_minv() below, which runs the unmodified reference on all-fp32 inputs.
    np.savez_compressed(os.path.join(OUT, "models.npz"), **d)
    print("models.npz", len(d))


def gen_minv():
    """use_Minv=True (ehf:183-184, 223-224, 331-332, 338-341, 415-417).  On the shipped dtypes (fp64 M and X) the
    reference raises "expected m1 and m2 to have the same dtype" at ehf:224: Minv = t.tensor(np.linalg.inv(M)) is
    fp64 and multiplies the fp32 AtXt buffer.  The UNMODIFIED reference does run the flag when every input is
    fp32 (M, X and the slice values): inv(M) is then an fp32 LAPACK inverse and the product an fp32 matmul.
    That all-fp32 run is what is stored here -- no shim, no edit; only the input dtypes differ from the
    experiments'.  T = 7, b = 3, un-normalised M (cond ~ 5), so the fp32 inverse is good to ~1e-6."""
    ns = load_ref_functions(3)
    T, N, F0, E = 7, 30, 3, 70
    C = random_coo(T, N, 0.1, 21)
    M = ns["create_matrix_M"](T, 3)
    Ct = ns["func_MProduct"](C, M)
    g = t.Generator().manual_seed(22)
    X = t.rand(T, N, F0, generator=g, dtype=t.float64)
    edges = t.stack([t.randint(0, T, (E,), generator=g), t.randint(0, N, (E,), generator=g),
                     t.randint(0, N, (E,), generator=g)])
    edges = edges[:, t.argsort(edges[0], stable=True)]
    dOut = t.randn(E, 2, generator=g)
    M32, X32 = M.float(), X.float()
    At32 = []
    for j in range(T):
        idx = Ct._indices()[0] == j
        At32.append(t.sparse.FloatTensor(Ct._indices()[1:3, idx], Ct._values()[idx].float()))
    d = {"Ct_idx": Ct._indices().numpy(), "Ct_val": Ct._values().numpy(), "M": M.numpy(), "X": X.numpy(),
         "edges": edges.numpy(), "dOut": dOut.numpy(), "TN": np.array([T, N])}
    t.manual_seed(800)
    m = ehf.EmbeddingGCN(At32, X32, edges, M32, hidden_feat=[5, 2], condensed_W=True, use_Minv=True)
    out = m()
    d["gcn1_W"], d["gcn1_U"], d["gcn1_out"] = m.W.detach().numpy().copy(), m.U.detach().numpy().copy(), out.detach().numpy()
    for k, v in grads(m, out, dOut, ["W", "U"]).items():
        d["gcn1_d" + k] = v
    t.manual_seed(810)
    m = ehf.EmbeddingGCN(At32, X32, edges, M32, hidden_feat=[5, 2], condensed_W=False, use_Minv=True)
    out = m()
    d["gcn1u_W"], d["gcn1u_U"], d["gcn1u_out"] = m.W.detach().numpy().copy(), m.U.detach().numpy().copy(), out.detach().numpy()
    for k, v in grads(m, out, dOut, ["W", "U"]).items():
        d["gcn1u_d" + k] = v
    t.manual_seed(820)
    try:
        m = ehf.EmbeddingGCN2(At32, X32, edges, M32, hidden_feat=[5, 4, 2], condensed_W=True, use_Minv=True,
                              nonlin2="leaky")
        out = m()
        for n in ("W1", "W2", "U"):
            d["gcn2_" + n] = getattr(m, n).detach().numpy().copy()
        d["gcn2_out"] = out.detach().numpy()
        for k, v in grads(m, out, dOut, ["W1", "W2", "U"]).items():
            d["gcn2_d" + k] = v
    except RuntimeError as ex:          # layer 2 of the reference up-casts to fp64 (ehf:335) before inv(M)
        d["gcn2_error"] = np.array(str(ex))
        print("EmbeddingGCN2(use_Minv=True) cannot run even in fp32:", ex)
    t.manual_seed(830)
    m = ehf.EmbeddingGCN_reg(At32, X32, M32, hidden_feat=[5, 2], condensed_W=True, use_Minv=True)
    out = m()
    dReg = t.randn(out.shape, generator=g)
    d["reg_W"], d["reg_lw"], d["reg_lb"] = (m.W.detach().numpy().copy(), m.lin1.weight.detach().numpy().copy(),
                                            m.lin1.bias.detach().numpy().copy())
    d["reg_out"], d["reg_dOut"] = out.detach().numpy(), dReg.numpy()
    for p_ in m.parameters():
        p_.grad = None
    out.backward(dReg)
    d["reg_dW"] = m.W.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "minv.npz"), **d)
    print("minv.npz", len(d))


def gen_preprocess():
    """read_data.py:88-188 helper functions executed from the reference text on a small random tensor."""
    src = open(os.path.join(REF, "read_data.py")).read().split("\n")

    def grab(name):
        a = [i for i, l in enumerate(src) if l.startswith("def " + name + "(")][0]
        b = a + 1
        while b < len(src) and (src[b].startswith((" ", "\t")) or src[b].strip() == ""):
            b += 1
        return "\n".join(src[a:b])
    ns = {"torch": t, "np": np, "edge_life_window": 3}
    for f in ("func_make_symmetric", "func_edge_life", "func_laplacian_transformation", "func_create_sparse"):
        exec(grab(f), ns)
    TT, N = 7, 30
    g = t.Generator().manual_seed(21)
    n = 260
    idx = t.stack([t.randint(0, TT, (n,), generator=g), t.randint(0, N, (n,), generator=g),
                   t.randint(0, N, (n,), generator=g)])
    A = t.sparse.DoubleTensor(idx, t.ones(n, dtype=t.double), t.Size([TT, N, N])).coalesce()
    B = ns["func_make_symmetric"](A, N, TT)
    E = ns["func_edge_life"](B, N, TT)
    C = ns["func_laplacian_transformation"](E, N, TT)
    S = ns["func_create_sparse"](C, N, TT, 4, 2, 6)
    d = {"TT_N_w": np.array([TT, N, 3])}
    for name, x in (("A", A), ("sym", B), ("life", E), ("lap", C), ("win", S)):
        d[name + "_idx"], d[name + "_val"] = x._indices().numpy(), x._values().numpy()
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **d)
    print("preprocess.npz", {k: v.shape for k, v in d.items() if k.endswith("_idx")})


def gen_data_helpers():
    """ehf:530-538 (compute_f1), 542-595 (load_data), 597-610 (create_node_features), 612-655 (split_data),
    669-729 (MRR / MAP) of the unmodified reference on a small random dataset in the MATLAB wire format of
    read_data.m:210-232 (``*_subs`` = nnz x 3, 1-based; ``*_vals`` = nnz x 1).

    Two environment shims, neither touches the reference's arithmetic: ``np.float`` (removed from numpy 1.24,
    used at ehf:677) is aliased to ``float``, and the subs are stored with an integer dtype because
    ``torch.Size`` no longer accepts the numpy floats MATLAB writes (ehf:549-550)."""
    import scipy.io as sio
    np.float = float
    ns = load_ref_functions(3)
    T, N, S_train, S_val, S_test = 10, 15, 6, 2, 2
    g = t.Generator().manual_seed(77)
    n = 170
    idx = t.stack([t.randint(0, T, (n,), generator=g), t.randint(0, N, (n,), generator=g),
                   t.randint(0, N, (n,), generator=g)])
    idx[:, 0] = t.tensor([T - 1, N - 1, N - 1])                      # pins the sizes load_data infers
    lab = t.randint(0, 2, (n,), generator=g).double() * 2 - 1          # chess-style +-1 labels
    A_labels = t.sparse_coo_tensor(idx, lab, (T, N, N)).coalesce()
    C = random_coo(T, N, 0.12, 78)
    M = ref_M_normalised(S_train, 3)

    def window(lo):
        sel = (C._indices()[0] >= lo) & (C._indices()[0] < lo + S_train)
        i = C._indices()[:, sel].clone()
        i[0] -= lo
        return t.sparse_coo_tensor(i, C._values()[sel], (S_train, N, N)).coalesce()
    Ct = {"train": ns["func_MProduct"](window(0), M), "val": ns["func_MProduct"](window(S_val), M),
          "test": ns["func_MProduct"](window(S_val + S_test), M)}

    def mat(sp):
        return sp._indices().t().numpy().astype(np.int64) + 1, sp._values().numpy().astype(np.float64)[:, None]
    m = {"M": M.numpy()}
    for name, sp in [("A_labels", A_labels), ("C", C)] + [("Ct_" + k, v) for k, v in Ct.items()]:
        m[name + "_subs"], m[name + "_vals"] = mat(sp)
    sio.savemat(os.path.join(OUT, "data_helpers.mat"), m, do_compression=True)

    d = {"sizes": np.array([T, N, S_train, S_val, S_test])}

    def put(name, sp):
        sp = sp.coalesce()
        d[name + "_idx"], d[name + "_val"], d[name + "_shape"] = sp._indices().numpy(), sp._values().numpy(), np.array(sp.shape)
    out = ehf.load_data(OUT + "/", "data_helpers.mat", S_train, S_val, S_test, True)
    put("tr_A", out[0]); put("tr_A_labels", out[1])
    for k, lst in zip(("train", "val", "test"), out[2:5]):
        for j, sl in enumerate(lst):
            put("tr_Ct_%s_%d" % (k, j), sl)
    d["tr_N"], d["tr_M"] = np.array(int(out[5])), out[6].numpy()
    out2 = ehf.load_data(OUT + "/", "data_helpers.mat", S_train, S_val, S_test, False)
    for k, lst in zip(("train", "val", "test"), out2[2:5]):
        d["raw_C_%s_len" % k] = np.array(len(lst))
        for j, sl in enumerate(lst):
            put("raw_C_%s_%d" % (k, j), sl)
    A = out[0]
    for sb in (True, False):
        for k, x in zip(("train", "val", "test"), ehf.create_node_features(A, S_train, S_val, S_test, sb)):
            d["X_%s_sb%d" % (k, sb)] = x.numpy()
    # edges = stored entries + random "negative" pairs, labels 0 (existing) / 1 (added), sorted by time as
    # augment_edges leaves them (ehf:518-525)
    edges = A_labels._indices()
    neg = t.stack([t.randint(0, T, (90,), generator=g), t.randint(0, N, (90,), generator=g),
                   t.randint(0, N, (90,), generator=g)])
    e_all = t.cat([edges, neg], 1)
    l_all = t.cat([t.zeros(edges.shape[1], dtype=t.long), t.ones(90, dtype=t.long)])
    _, order = e_all[0].sort()
    e_all, l_all = e_all[:, order], l_all[order]
    d["edges_aug"], d["labels"] = e_all.numpy(), l_all.numpy()
    names_sb = ["edges_train", "target_train", "e_train", "edges_val", "target_val", "e_val", "K_val", "edges_test",
                "target_test", "e_test", "K_test"]
    names = [x for x in names_sb if not x.startswith("K_")]
    for sb, nm in ((True, names_sb), (False, names)):
        res = ehf.split_data(e_all.clone(), l_all.clone(), S_train, S_val, S_test, sb)
        for k, x in zip(nm, res):
            d["split_sb%d_%s" % (sb, k)] = np.asarray(x)
    guess = t.randint(0, 2, (300,), generator=g)
    target = t.randint(0, 2, (300,), generator=g)
    d["f1_guess"], d["f1_target"] = guess.numpy(), target.numpy()
    d["f1_out"] = np.array([float(x) for x in ehf.compute_f1(guess, target)])
    for dt, tag in ((t.float32, "f32"), (t.float64, "f64")):
        logits = t.randn(e_all.shape[1], 2, generator=g, dtype=t.float64).to(dt)
        MAP, MRR = ehf.compute_MAP_MRR(logits, l_all, e_all)
        d["metric_logits_" + tag] = logits.numpy()
        d["metric_out_" + tag] = np.array([float(MAP), float(MRR)])
    np.savez_compressed(os.path.join(OUT, "data_helpers.npz"), **d)
    print("data_helpers.npz", len(d), "MAP/MRR", d["metric_out_f32"], d["metric_out_f64"], "f1", d["f1_out"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "minv":       # only the newest fixture (the others are unchanged)
        gen_minv()
        sys.exit(0)
    gen_minv()
    gen_data_helpers()
    gen_preprocess()
    gen_mproduct()
    gen_chess()
    gen_models()
