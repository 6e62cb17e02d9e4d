"""the .mat wire format, node features, data split and metrics around the path.

Golden = outputs of the reference's own load_data / create_node_features / split_data / compute_f1 /
compute_MAP_MRR (tests/golden/make_golden.py::gen_data_helpers, data_helpers.mat + data_helpers.npz).
CPU tests pin the oracle restatement and the device-agnostic host logic; the gpu tests run the same functions
on the device and the loader through the C ABI.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

SPLIT_NAMES = ["edges_train", "target_train", "e_train", "edges_val", "target_val", "e_val", "K_val", "edges_test",
               "target_test", "e_test", "K_test"]


def _sizes(g):
    return tuple(int(x) for x in g["sizes"])


def _split_names(sb):
    return SPLIT_NAMES if sb else [n for n in SPLIT_NAMES if not n.startswith("K_")]


def _random_metric_case(seed, E=4000, T=6, N=40, dtype=torch.float64):
    """edges with duplicates, rows without label-1 cells and negative as well as positive logits"""
    g = torch.Generator().manual_seed(seed)
    edges = torch.stack([torch.randint(0, T, (E,), generator=g), torch.randint(0, N, (E,), generator=g),
                         torch.randint(0, N, (E,), generator=g)])
    target = (torch.rand(E, generator=g) < 0.3).long()
    logits = torch.randn(E, 2, generator=g, dtype=torch.float64).to(dtype)
    return logits, target, edges


# ------------------------------------------------------------------ oracle vs the reference's outputs
def test_oracle_f1_and_metrics(golden_data):
    g = golden_data
    np.testing.assert_allclose(oracle.compute_f1(g["f1_guess"], g["f1_target"]), g["f1_out"], rtol=1e-15)
    for tag in ("f32", "f64"):
        got = oracle.compute_MAP_MRR(g["metric_logits_" + tag], g["labels"], g["edges_aug"])
        np.testing.assert_allclose(got, g["metric_out_" + tag], rtol=1e-13)


def test_oracle_features_and_split(golden_data):
    g = golden_data
    T, N, S_train, S_val, S_test = _sizes(g)
    for sb in (True, False):
        X = oracle.create_node_features(g["tr_A_idx"], g["tr_A_val"], T, N, S_train, S_val, S_test, sb)
        for x, k in zip(X, ("train", "val", "test")):
            assert np.array_equal(x, g["X_%s_sb%d" % (k, sb)])
        res = oracle.split_data(g["edges_aug"], g["labels"], S_train, S_val, S_test, sb)
        for r, k in zip(res, _split_names(sb)):
            assert np.array_equal(np.asarray(r), g["split_sb%d_%s" % (sb, k)]), k


def test_oracle_mat_loader(golden_data):
    g = golden_data
    T, N, S_train, S_val, S_test = _sizes(g)
    path = os.path.join(GOLDEN, "data_helpers.mat")
    (a_idx, a_val), blocks, n, M = oracle.load_data_mat(path, S_train, S_val, S_test, True)
    assert n == int(g["tr_N"]) and np.array_equal(a_idx, g["tr_A_labels_idx"]) and np.array_equal(a_val, g["tr_A_labels_val"])
    assert np.array_equal(M, g["tr_M"])
    for k, blk in zip(("train", "val", "test"), blocks):
        for j, (idx, val) in enumerate(blk):
            assert np.array_equal(idx, g["tr_Ct_%s_%d_idx" % (k, j)])
            assert np.array_equal(val, g["tr_Ct_%s_%d_val" % (k, j)])
    _, blocks, _ = oracle.load_data_mat(path, S_train, S_val, S_test, False)
    for k, blk in zip(("train", "val", "test"), blocks):
        assert len(blk) == int(g["raw_C_%s_len" % k])
        for j, (idx, val) in enumerate(blk):
            assert np.array_equal(idx, g["raw_C_%s_%d_idx" % (k, j)])
            assert np.array_equal(val, g["raw_C_%s_%d_val" % (k, j)])


# ------------------------------------------------------------------ host logic of the product (any device)
def _check_product_against_golden(g, dev):
    from tmgcn_b200 import data
    T, N, S_train, S_val, S_test = _sizes(g)
    f1 = data.compute_f1(torch.from_numpy(g["f1_guess"]).to(dev), torch.from_numpy(g["f1_target"]).to(dev))
    np.testing.assert_allclose([float(x) for x in f1], g["f1_out"], rtol=1e-15)
    for tag in ("f32", "f64"):
        out = data.compute_MAP_MRR(torch.from_numpy(g["metric_logits_" + tag]).to(dev),
                                   torch.from_numpy(g["labels"]).to(dev), torch.from_numpy(g["edges_aug"]).to(dev))
        assert all(o.dtype == torch.float64 for o in out)
        # fp32 scores: the device softmax may differ from the CPU's in the last bit, which can reorder near-ties
        tol = 1e-12 if (tag == "f64" or dev.type == "cpu") else 1e-5
        np.testing.assert_allclose([float(x) for x in out], g["metric_out_" + tag], rtol=tol)
    A = torch.sparse_coo_tensor(torch.from_numpy(g["tr_A_idx"]), torch.from_numpy(g["tr_A_val"]), (T, N, N)).to(dev)
    for sb in (True, False):
        X = data.create_node_features(A, S_train, S_val, S_test, sb)
        for x, k in zip(X, ("train", "val", "test")):
            assert x.dtype == torch.float64 and np.array_equal(x.cpu().numpy(), g["X_%s_sb%d" % (k, sb)])
        e_in, l_in = torch.from_numpy(g["edges_aug"]).to(dev), torch.from_numpy(g["labels"]).to(dev)
        res = data.split_data(e_in, l_in, S_train, S_val, S_test, sb)
        assert len(res) == len(_split_names(sb))
        for r, k in zip(res, _split_names(sb)):
            assert np.array_equal(r.cpu().numpy(), g["split_sb%d_%s" % (sb, k)]), k
        assert np.array_equal(e_in.cpu().numpy(), g["edges_aug"])        # the input is left alone


def test_product_helpers_cpu(golden_data):
    _check_product_against_golden(golden_data, torch.device("cpu"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_product_metrics_vs_oracle_cpu(dtype):
    from tmgcn_b200 import data
    logits, target, edges = _random_metric_case(5, dtype=dtype)
    want = oracle.compute_MAP_MRR(logits.numpy(), target.numpy(), edges.numpy())
    got = data.compute_MAP_MRR(logits, target, edges)
    np.testing.assert_allclose([float(x) for x in got], want, rtol=1e-11)
    # one slice alone, through the single-slice entry points of ehf:684-711
    m = edges[0] == 2
    ap = data.get_MAP(logits[m], target[m])
    assert abs(float(ap) - oracle.average_precision_class0(target[m].numpy(),
                                                           torch.softmax(logits[m], 1)[:, 0].numpy())) < 1e-12


def test_mrr_is_nan_without_negative_rows():
    """a slice whose rows hold no label-1 cell has an empty mean (ehf:701) -> NaN, like the reference"""
    from tmgcn_b200 import data
    edges = torch.tensor([[0, 0, 1, 1], [0, 1, 0, 2], [1, 2, 2, 0]])
    target = torch.tensor([0, 0, 0, 1])
    logits = torch.tensor([[0.3, 0.1], [-0.2, 0.4], [0.5, 0.0], [0.1, 0.2]], dtype=torch.float64)
    MAP, MRR = data.compute_MAP_MRR(logits, target, edges)
    want = oracle.compute_MAP_MRR(logits.numpy(), target.numpy(), edges.numpy())
    assert torch.isnan(MRR) and np.isnan(want[1])
    np.testing.assert_allclose(float(MAP), want[0], rtol=1e-14)


def test_save_mat_writes_the_matlab_layout(tmp_path, golden_data):
    import scipy.io as sio
    from tmgcn_b200 import data
    g = golden_data
    T, N = _sizes(g)[:2]
    A = torch.sparse_coo_tensor(torch.from_numpy(g["tr_A_labels_idx"]), torch.from_numpy(g["tr_A_labels_val"]), (T, N, N))
    data.save_mat(str(tmp_path / "x.mat"), M=torch.from_numpy(g["tr_M"]), A_labels=A)
    ref = sio.loadmat(os.path.join(GOLDEN, "data_helpers.mat"))
    got = sio.loadmat(str(tmp_path / "x.mat"))
    assert got["A_labels_subs"].dtype == np.float64 and got["A_labels_subs"].shape[1] == 3   # 1-based doubles
    assert np.array_equal(got["A_labels_subs"], ref["A_labels_subs"].astype(np.float64))
    assert np.array_equal(got["A_labels_vals"], ref["A_labels_vals"])
    assert np.array_equal(got["M"], ref["M"])


# ------------------------------------------------------------------ on the device
@pytest.mark.gpu
def test_product_helpers_gpu(golden_data):
    _check_product_against_golden(golden_data, torch.device("cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_product_metrics_vs_oracle_gpu(dtype):
    from tmgcn_b200 import data
    logits, target, edges = _random_metric_case(9, E=20000, T=9, N=70, dtype=dtype)
    want = oracle.compute_MAP_MRR(logits.numpy(), target.numpy(), edges.numpy())
    got = data.compute_MAP_MRR(logits.cuda(), target.cuda(), edges.cuda())
    assert got[0].is_cuda and got[1].is_cuda
    np.testing.assert_allclose([float(x) for x in got], want, rtol=1e-11 if dtype == torch.float64 else 1e-5)


def _csr_slices(csr):
    idx, val = csr.to_coo()
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    return [(idx[1:3, idx[0] == j], val[idx[0] == j]) for j in range(csr.T)]


@pytest.mark.gpu
def test_load_data_matches_reference(golden_data):
    from tmgcn_b200 import SliceCSR, data
    g = golden_data
    T, N, S_train, S_val, S_test = _sizes(g)
    out = data.load_data(GOLDEN, "data_helpers.mat", S_train, S_val, S_test, True, dtype=torch.float64)
    A, A_labels, blocks, n, M = out[0], out[1], out[2:5], out[5], out[6]
    assert n == int(g["tr_N"]) and tuple(A.shape) == tuple(g["tr_A_shape"]) and A.is_cuda
    assert A.dtype == torch.float32 and A_labels.dtype == torch.float64 and M.dtype == torch.float64
    assert np.array_equal(A._indices().cpu().numpy(), g["tr_A_idx"]) and np.array_equal(A._values().cpu().numpy(), g["tr_A_val"])
    assert np.array_equal(A_labels._indices().cpu().numpy(), g["tr_A_labels_idx"])
    assert np.array_equal(A_labels._values().cpu().numpy(), g["tr_A_labels_val"])
    assert np.array_equal(M.numpy(), g["tr_M"])
    for k, csr in zip(("train", "val", "test"), blocks):
        assert isinstance(csr, SliceCSR) and csr.T == S_train and csr.N == N
        for j, (idx, val) in enumerate(_csr_slices(csr)):
            assert np.array_equal(idx, g["tr_Ct_%s_%d_idx" % (k, j)])            # bit-exact indices
            assert np.array_equal(val, g["tr_Ct_%s_%d_val" % (k, j)])            # fp64 values untouched
    out = data.load_data(GOLDEN, "data_helpers.mat", S_train, S_val, S_test, False, dtype=torch.float64)
    for k, csr in zip(("train", "val", "test"), out[2:5]):
        assert csr.T == int(g["raw_C_%s_len" % k])
        for j, (idx, val) in enumerate(_csr_slices(csr)):
            assert np.array_equal(idx, g["raw_C_%s_%d_idx" % (k, j)])
            assert np.array_equal(val, g["raw_C_%s_%d_val" % (k, j)])
    # the reference's list-of-sparse-matrices form, and the default fp32 layout the kernels consume
    lst = data.load_data(GOLDEN, "data_helpers.mat", S_train, S_val, S_test, True, as_list=True)[2]
    assert len(lst) == S_train and lst[0].layout == torch.sparse_coo and lst[0].dtype == torch.float32
    np.testing.assert_allclose(lst[3]._values().cpu().numpy(), g["tr_Ct_train_3_val"], rtol=1e-7)


@pytest.mark.gpu
def test_mat_round_trip_and_model_run(tmp_path, golden_data):
    """save_mat -> load_data -> features -> split -> EmbeddingGCN -> metrics: the chain the experiment scripts run
    (experiment_chess_our_link_prediction.py:40-108) end to end on the device"""
    import tmgcn_b200 as tg
    from tmgcn_b200 import data
    g = golden_data
    T, N, S_train, S_val, S_test = _sizes(g)
    first = data.load_data(GOLDEN, "data_helpers.mat", S_train, S_val, S_test, True, dtype=torch.float64)
    data.save_mat(str(tmp_path / "again.mat"), M=first[6], A_labels=first[1], Ct_train=first[2], Ct_val=first[3],
                  Ct_test=first[4])
    A, A_labels, Ct_train, Ct_val, Ct_test, n, M = data.load_data(str(tmp_path), "again.mat", S_train, S_val, S_test, True)
    assert torch.equal(Ct_train.col, first[2].col) and torch.equal(Ct_train.rowptr, first[2].rowptr)
    X_train, X_val, X_test = data.create_node_features(A, S_train, S_val, S_test, True)
    edges = torch.from_numpy(g["edges_aug"]).cuda()
    labels = torch.from_numpy(g["labels"]).cuda()
    edges_train, target_train, e_train, edges_val, target_val, e_val, K_val, *_ = data.split_data(
        edges, labels, S_train, S_val, S_test, True)
    torch.manual_seed(0)
    gcn = tg.EmbeddingGCN(Ct_train, X_train, e_train, M, hidden_feat=[6, 2], condensed_W=True, use_Minv=False)
    out_train = gcn()
    out_val = gcn(Ct_val, X_val, e_val)
    assert out_train.shape == (e_train.shape[1], 2) and out_val.shape == (e_val.shape[1], 2)
    guess = out_train.argmax(dim=1)
    tt = target_train[edges_train[0] != 0]
    p, r, f = data.compute_f1(guess, tt)
    MAP, MRR = data.compute_MAP_MRR(out_train, tt, edges_train[:, edges_train[0] != 0])
    assert 0.0 <= float(MAP) <= 1.0 and 0.0 <= float(MRR) <= 1.0 and torch.isfinite(out_val).all()
    want = oracle.compute_MAP_MRR(out_train.detach().cpu().numpy(), tt.cpu().numpy(),
                                  edges_train[:, edges_train[0] != 0].cpu().numpy())
    np.testing.assert_allclose([float(MAP), float(MRR)], want, rtol=1e-5)         # fp32 logits
