"""Device-resident input pipeline: fm_for_online_recommendation_b200/data.py + csrc/dataset.cu against the
reference's utils/data_preprocess.py.  Everything here is integer / byte work: the bar is bit-exact.

CPU tests pin the oracle (oracle/dataset.py) on tests/golden/dataset.npz, which tests/golden/make_dataset_golden.py generated
from the reference's own functions; GPU tests compare the device pipeline with the fixture (same files, same `random` seeds)
and, at cod-rna's full size (59 535 x 8) and beyond, with the oracle."""
import importlib
import os
import random
import sys

import numpy as np
import pytest

from _util import ROOT, load_golden

from oracle import dataset as orc


# ----------------------------------------------------------------------------- oracle pinned on the reference (CPU)
def test_oracle_matches_reference_fixture():
    g = load_golden("dataset")
    codes, sizes = orc.first_seen_encode_loop(g["svm_X"])
    assert np.array_equal(codes, g["svm_index"]) and np.array_equal(sizes, g["svm_sizes"])
    codes2, sizes2 = orc.first_seen_encode(g["svm_X"])
    assert np.array_equal(codes2, codes) and np.array_equal(sizes2, sizes)
    random.seed(7)
    lists, ratio = orc.create_ten_iter_indices(g["cr_label"], 5, 40)
    assert np.array_equal(g["cr_index"][np.asarray(lists)], g["ti_Xi"]) and np.array_equal(np.asarray(ratio), g["ti_ratio"])
    assert np.array_equal(g["cr_label"][np.asarray(lists)], g["ti_Y"])
    random.seed(8)
    lists, ratio = orc.create_dataset_indices(g["cr_label"], 2, 5, 40)
    assert np.array_equal(g["cr_index"][np.asarray(lists)], g["cd_Xi"]) and np.array_equal(np.asarray(ratio), g["cd_ratio"])
    random.seed(9)
    ix = orc.balance_indices(g["cr_label"])
    assert np.array_equal(g["cr_index"][ix], g["bc_index"]) and np.array_equal(g["cr_label"][ix], g["bc_label"])
    random.seed(10)
    lab = np.where(g["svm_y"] == -1, 0, g["svm_y"])
    ix = orc.balance_indices(lab)
    assert np.array_equal(g["svm_index"][ix], g["bs_index"]) and np.array_equal(g["svm_value"][ix], g["bs_value"])


def test_oracle_fast_encode_equals_the_loop_on_edge_cases():
    rs = np.random.RandomState(3)
    X = rs.randint(-3, 4, size=(300, 5)).astype(np.float64)
    X[rs.rand(300, 5) < 0.2] = -0.0                      # -0.0 == 0.0 for list.index
    X[:, 4] = 2.5                                        # a constant column
    a, sa = orc.first_seen_encode_loop(X)
    b, sb = orc.first_seen_encode(X)
    assert np.array_equal(a, b) and np.array_equal(sa, sb) and sa[4] == 1


# ----------------------------------------------------------------------------- device pipeline (GPU)
def _write_criteo(tmp_path, g):
    data, emb = os.path.join(tmp_path, "tiny_train_input.csv"), os.path.join(tmp_path, "category_emb.csv")
    with open(data, "w") as f:
        for lab, row in zip(g["cr_label"], g["cr_index"]):
            f.write(",".join([str(int(lab))] + [str(int(v)) for v in row]) + "\n")
    with open(emb, "w") as f:
        for fld, s in enumerate(g["cr_sizes"]):
            for c in range(int(s)):
                f.write(f"{fld},c{c},{c}\n")
    return data, emb


def _local(e, sizes):
    off = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])[:-1]
    return e.ids.cpu().numpy().astype(np.int64) - off[None, :]


@pytest.mark.gpu
def test_criteo_builders_equal_the_reference(tmp_path):
    from fm_for_online_recommendation_b200 import data
    g = load_golden("dataset")
    fdata, femb = _write_criteo(str(tmp_path), g)
    ds = data.read_criteo_data(fdata, femb)
    sizes = g["cr_sizes"]
    assert list(ds.feature_sizes) == list(sizes) and len(ds) == 600 and ds.xv is None
    assert np.array_equal(_local(ds.batch(0, 600), sizes), g["cr_read_index"])
    assert np.array_equal(ds.y.cpu().numpy(), g["cr_read_label"].astype(np.float32))
    Xi, Xv, Y, ratio = data.construct_batch_criteo_data(ds, 100, 5)
    for i in range(5):
        assert np.array_equal(_local(Xi[i], sizes), g["cb_Xi"][i]) and Y[i] == g["cb_Y"][i].tolist()
        assert Xi[i].xv is None and np.all(g["cb_Xv"][i] == 1)           # all-ones values are implicit
        assert Xi[i].ids.data_ptr() == ds.ids[100 * i:].data_ptr()       # a view, not a copy
    assert np.array_equal(np.asarray(ratio), g["cb_ratio"])
    random.seed(7)
    Xi, Xv, Y, ratio = data.create_ten_iter(ds, 5, 40)
    for i in range(5):
        assert np.array_equal(_local(Xi[i], sizes), g["ti_Xi"][i]) and Y[i] == g["ti_Y"][i].tolist()
        assert np.array_equal(Xi[i].y.cpu().numpy(), g["ti_Y"][i].astype(np.float32))
    assert np.array_equal(np.asarray(ratio), g["ti_ratio"])
    random.seed(8)
    Xi, Xv, Y, ratio = data.create_dataset(ds, 2, 5, 40)
    for i in range(5):
        assert np.array_equal(_local(Xi[i], sizes), g["cd_Xi"][i]) and Y[i] == g["cd_Y"][i].tolist()
    assert np.array_equal(np.asarray(ratio), g["cd_ratio"])
    random.seed(9)
    b = data.balance(ds)
    assert np.array_equal(_local(b.batch(0, len(b)), sizes), g["bc_index"])
    assert np.array_equal(b.y.cpu().numpy(), g["bc_label"].astype(np.float32))
    with pytest.raises(IndexError):
        data.construct_batch_criteo_data(ds, 100, 7)                     # the reference's list index runs off the end
    with pytest.raises(IndexError):
        ds.take([0, 600])
    with pytest.raises(IndexError):
        data.DeviceDataset.from_local_ids([[0] * 38 + [int(sizes[38])]], None, [1], sizes)


@pytest.mark.gpu
def test_svm_reader_equals_the_reference(tmp_path):
    from fm_for_online_recommendation_b200 import data
    g = load_golden("dataset")
    path = os.path.join(str(tmp_path), "cod-rna")
    with open(path, "w") as f:
        for yy, row in zip(g["svm_y"], g["svm_X"]):
            f.write(str(int(yy)) + " " + " ".join(f"{c + 1}:{float(row[c])!r}" for c in range(8) if row[c] != 0.0) + "\n")
    ds = data.read_svm_file(path)
    assert np.array_equal(np.asarray(ds.feature_sizes), g["svm_sizes"])
    assert np.array_equal(_local(ds.batch(0, len(ds)), g["svm_sizes"]), g["svm_index"])
    assert np.array_equal(ds.y.cpu().numpy(), g["svm_label"].astype(np.float32))
    assert np.array_equal(ds.xv.cpu().numpy(), g["svm_value"].astype(np.float32))
    random.seed(10)
    b = data.balance(ds)
    assert np.array_equal(_local(b.batch(0, len(b)), g["svm_sizes"]), g["bs_index"])
    assert np.array_equal(b.xv.cpu().numpy(), g["bs_value"].astype(np.float32))
    assert np.array_equal(b.y.cpu().numpy(), g["bs_label"].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("N,d,pool", [(1, 3, 1), (33, 1, 4), (4097, 8, 50), (59535, 8, 3000), (300000, 5, 200000)])
def test_dict_encode_bit_exact_vs_oracle(N, d, pool):
    """cod-rna's full size (BASELINE.json configs[1]: 59 535 x 8) and a column with more distinct values than one scan round"""
    from fm_for_online_recommendation_b200 import data
    rs = np.random.RandomState(N + d)
    vals = rs.randn(pool)
    X = vals[rs.randint(0, pool, size=(N, d))]
    X[rs.rand(N, d) < 0.05] = -0.0
    X[rs.rand(N, d) < 0.05] = 0.0
    if d > 1:
        X[:, d - 1] = 7.0
    codes, sizes = data.dict_encode_first_seen(X)
    oc, osz = (orc.first_seen_encode_loop if N <= 4097 else orc.first_seen_encode)(X)
    assert np.array_equal(sizes, osz)
    assert np.array_equal(codes.cpu().numpy().astype(np.int64), oc)
    # size-independent properties: codes of a column are a bijection onto range(size), first appearances ascend
    c0 = codes[:, 0].cpu().numpy()
    first = np.full(int(sizes[0]), N, np.int64)
    np.minimum.at(first, c0, np.arange(N))
    assert np.all(np.diff(first) > 0) and len(np.unique(c0)) == sizes[0]
    with pytest.raises(ValueError):
        data.dict_encode_first_seen(np.array([[1.0, np.nan]]))


@pytest.mark.gpu
def test_take_full_size_round_trip_and_model_accepts_the_batches():
    """a Criteo-shaped data set (39 fields, main_experiment.py:56-58 sizes): take(perm) then take(inverse perm) is the
    identity; the positives count matches; the model's update_embedding on a pipeline batch equals the list path bit for bit"""
    import torch
    from fm_for_online_recommendation_b200 import FMAdam, data
    from test_gpu_main_experiment import feature_sizes
    from _util import synth
    n = 200000
    Xi, Xv, Y = synth(feature_sizes, n, 5)
    ds = data.DeviceDataset.from_local_ids(Xi, Xv, Y, feature_sizes)
    assert ds.xv is None
    rs = np.random.RandomState(1)
    perm = rs.permutation(n)
    e, pos = ds.take(perm, count_positives=True)
    assert pos == int((Y == 1).sum())
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    back = data.DeviceDataset(e.ids, e.xv, e.y, feature_sizes).take(torch.from_numpy(inv).cuda())
    assert torch.equal(back.ids, ds.ids) and torch.equal(back.y, ds.y)
    assert np.array_equal(_local(e, feature_sizes)[:1000], Xi[perm[:1000]])
    # the same step from lists and from pipeline batches
    torch.manual_seed(0)
    m1 = FMAdam(feature_sizes, embedding_size=10, n=0.001)
    torch.manual_seed(0)
    m2 = FMAdam(feature_sizes, embedding_size=10, n=0.001)
    Xb, _, Yb, _ = data.construct_batch_criteo_data(ds, 2500, 3)
    for i in range(3):
        lo = 2500 * i
        l1 = m1.update_embedding(Xi[lo:lo + 2500].tolist(), Xv[lo:lo + 2500].tolist(), [int(v) for v in Y[lo:lo + 2500]])
        l2 = m2.update_embedding(Xb[i], Xb[i], Yb[i])
        assert float(l1) == float(l2)
    assert torch.equal(m1._table, m2._table)
    other = FMAdam(feature_sizes[:-1] + [feature_sizes[-1] + 1], embedding_size=10)
    with pytest.raises(IndexError):
        other.update_embedding(Xb[0], Xb[0], Yb[0])          # encoded for another table layout


@pytest.mark.gpu
def test_dropin_data_preprocess_module(tmp_path):
    """`from utils import data_preprocess` (main_experiment.py:1,26-36) through the drop-in path"""
    g = load_golden("dataset")
    fdata, femb = _write_criteo(str(tmp_path), g)
    dropin = os.path.join(ROOT, "fm_for_online_recommendation_b200", "dropin")
    saved = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, dropin)
    try:
        dp = importlib.import_module("utils.data_preprocess")
        train_dict = dp.read_criteo_data(fdata, femb)
        assert train_dict["size"] == 600 and train_dict["feature_sizes"] == g["cr_sizes"].tolist()
        random.seed(7)
        Xi, Xv, Y, ratio = dp.create_ten_iter(fdata, femb, 5, 40)
        assert [tuple(r) for r in ratio] == [tuple(r) for r in g["ti_ratio"].tolist()]
        assert np.array_equal(_local(Xi[4], g["cr_sizes"]), g["ti_Xi"][4]) and Y[4] == g["ti_Y"][4].tolist()
        from fm_for_online_recommendation_b200 import FMAdam
        m = FMAdam(train_dict["feature_sizes"], embedding_size=4, n=0.01)
        loss = m.update_embedding(Xi[2], Xv[2], Y[2])
        pred = m.predict(Xi[2], Xv[2])
        assert np.isfinite(float(loss.cpu().data)) and pred.shape == (40,)
        t, acc, roc, conf = m.run_experiment(Xi[0], Xv[0], Y[0])
        assert sum(conf.values()) == 40
    finally:
        sys.path.remove(dropin)
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
