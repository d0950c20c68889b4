"""Oracle (test infrastructure, CPU) for the device-resident input pipeline: numpy / pure-Python restatement of the
reference's utils/data_preprocess.py.  Only tests/ import this; the product never does.

Pinned by tests/golden/dataset.npz, generated from the reference's own functions by tests/golden/make_dataset_golden.py
(tests/test_dataset.py::test_oracle_matches_reference_fixture).
"""
import random

import numpy as np


def first_seen_encode_loop(X):
    """data_preprocess.py:100-108 verbatim in meaning: per cell, the value's index in the column's list of values seen so
    far, appended on a miss (`list.index` compares with ==)."""
    X = np.asarray(X, dtype=np.float64)
    vocab = [[] for _ in range(X.shape[1])]
    codes = np.zeros(X.shape, dtype=np.int64)
    for r in range(X.shape[0]):
        for c in range(X.shape[1]):
            v = float(X[r, c])
            try:
                codes[r, c] = vocab[c].index(v)
            except ValueError:
                vocab[c].append(v)
                codes[r, c] = len(vocab[c]) - 1
    return codes, np.asarray([len(v) for v in vocab], dtype=np.int64)


def first_seen_encode(X):
    """same result in O(N log N): rank of the value's first row among the first rows of the column's distinct values"""
    X = np.asarray(X, dtype=np.float64) + 0.0          # -0.0 + 0.0 = +0.0: one key, as == sees them
    codes = np.zeros(X.shape, dtype=np.int64)
    sizes = np.zeros(X.shape[1], dtype=np.int64)
    for c in range(X.shape[1]):
        _, first, inv = np.unique(X[:, c], return_index=True, return_inverse=True)
        rank = np.empty(first.size, dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(first.size)
        codes[:, c] = rank[inv.reshape(-1)]
        sizes[c] = first.size
    return codes, sizes


def split_by_label(label):
    """_find_pos_and_neg (:183-190)"""
    label = np.asarray(label)
    return {"0": np.nonzero(label == 0)[0].tolist(), "1": np.nonzero(label == 1)[0].tolist()}


def create_ten_iter_indices(label, num_batch, num_batchdata, rng=random):
    """create_ten_iter (:193-229): the row numbers of every batch and the ratio list"""
    ratios = split_by_label(label)
    out, ratio_list = [], []
    for i in range(num_batch):
        num_pos = int(num_batchdata / num_batch * (i + 1))
        num_neg = num_batchdata - num_pos
        ratio_list.append((num_neg, num_pos))
        ind = ratios["1"][:num_pos] + ratios["0"][:num_neg]
        ratios["1"], ratios["0"] = ratios["1"][num_pos:], ratios["0"][num_neg:]
        rng.shuffle(ind)
        out.append(ind)
    return out, ratio_list


def create_dataset_indices(label, batch_ratio, num_batch, num_batchdata, rng=random):
    """create_dataset (:232-262)"""
    ratios = split_by_label(label)
    out, ratio_list = [], []
    for _ in range(num_batch):
        ratio_list.append((batch_ratio, num_batch - batch_ratio))
        num_pos = int(num_batchdata / num_batch * batch_ratio)
        num_neg = num_batchdata - num_pos
        ind = ratios["1"][:num_pos] + ratios["0"][:num_neg]
        ratios["1"], ratios["0"] = ratios["1"][num_pos:], ratios["0"][num_neg:]
        rng.shuffle(ind)
        out.append(ind)
    return out, ratio_list


def balance_indices(label, rng=random):
    """balance_criteo_data / balance_svm_data (:46-82, :120-151)"""
    idx = split_by_label(label)
    rng.shuffle(idx["0"])
    idx["0"] = idx["0"][:len(idx["1"])]
    idx["0"].extend(idx["1"])
    rng.shuffle(idx["0"])
    return idx["0"]
