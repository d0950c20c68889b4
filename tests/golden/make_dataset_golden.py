"""Golden fixture for the input pipeline, generated from the REFERENCE's utils/data_preprocess.py (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_dataset_golden.py

Writes small synthetic files in the reference's formats (Criteo-tiny csv + category file, svmlight with 8 features) to a
temporary directory, runs read_criteo_data / _construct_batch_criteo_data / create_ten_iter / create_dataset /
balance_criteo_data / read_svm_file / balance_svm_data on them with fixed `random` seeds, and stores the raw inputs and every
output in tests/golden/dataset.npz.  /root/reference does not exist on the GPU box: tests read only the .npz.
"""
import os
import random
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from utils import data_preprocess as dp  # noqa: E402


def write_criteo(dirname, rs, n=600):
    sizes = rs.randint(2, 40, size=39)
    sizes[9], sizes[21] = 3, 2
    idx = np.stack([rs.randint(0, s, size=n) for s in sizes], axis=1)
    label = (rs.rand(n) < 0.45).astype(np.int64)
    data = os.path.join(dirname, "tiny_train_input.csv")
    emb = os.path.join(dirname, "category_emb.csv")
    with open(data, "w") as f:
        for i in range(n):
            f.write(",".join([str(label[i])] + [str(v) for v in idx[i]]) + "\n")
    with open(emb, "w") as f:
        for fld, s in enumerate(sizes):
            for c in range(s):
                f.write(f"{fld},c{c},{c}\n")
    return data, emb, sizes, idx, label


def write_svm(dirname, rs, n=400):
    pools = [np.round(rs.randn(rs.randint(3, 60)), 3) for _ in range(8)]
    X = np.stack([p[rs.randint(0, p.size, size=n)] for p in pools], axis=1)
    X[rs.rand(n, 8) < 0.1] = 0.0
    y = np.where(rs.rand(n) < 0.4, 1, -1)
    path = os.path.join(dirname, "cod-rna")
    with open(path, "w") as f:
        for i in range(n):
            f.write(str(y[i]) + " " + " ".join(f"{c + 1}:{float(X[i, c])!r}" for c in range(8) if X[i, c] != 0.0) + "\n")
    return path, X, y


def main():
    rs = np.random.RandomState(11)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        data, emb, sizes, idx, label = write_criteo(d, rs)
        out.update(cr_sizes=sizes, cr_index=idx, cr_label=label)
        res = dp.read_criteo_data(data, emb)
        assert res["feature_sizes"] == list(sizes) and res["size"] == len(label)
        out["cr_read_index"] = np.asarray(res["index"])
        out["cr_read_label"] = np.asarray(res["label"])
        Xi, Xv, Y, ratio = dp._construct_batch_criteo_data(res, 100, 5)
        out.update(cb_Xi=np.asarray(Xi), cb_Xv=np.asarray(Xv), cb_Y=np.asarray(Y), cb_ratio=np.asarray(ratio))
        random.seed(7)
        Xi, Xv, Y, ratio = dp.create_ten_iter(data, emb, 5, 40)
        out.update(ti_Xi=np.asarray(Xi), ti_Y=np.asarray(Y), ti_ratio=np.asarray(ratio))
        random.seed(8)
        Xi, Xv, Y, ratio = dp.create_dataset(data, emb, 2, 5, 40)
        out.update(cd_Xi=np.asarray(Xi), cd_Y=np.asarray(Y), cd_ratio=np.asarray(ratio))
        random.seed(9)
        res = dp.balance_criteo_data(data, emb)
        out.update(bc_index=np.asarray(res["index"]), bc_label=np.asarray(res["label"]))

        path, X, y = write_svm(d, rs)
        out.update(svm_X=X, svm_y=y)
        res = dp.read_svm_file(path)
        assert np.array_equal(res["value"], X)
        out.update(svm_index=np.asarray(res["index"]), svm_label=np.asarray(res["label"]),
                   svm_sizes=np.asarray(res["feature_sizes"]), svm_value=np.asarray(res["value"]))
        random.seed(10)
        res = dp.balance_svm_data(path)
        out.update(bs_index=np.asarray(res["index"]), bs_label=np.asarray(res["label"]), bs_value=np.asarray(res["value"]))
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
