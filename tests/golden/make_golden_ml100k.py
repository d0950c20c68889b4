"""Known answers on the reference's BUNDLED ml-100k data (SURVEY.md section 4's table; build container only).

`dataset/ml-100k/ua.base`, ordered by `np.argsort(timestamp.astype(float32), kind='stable')` (what
utils/data_manager.py:51-60 does), first 20 000 ratings, one-hot user | item | bias -> d = 2626 (data_manager.py:18-48),
task 'reg', eta = 0.005, m = 5; the three classical learners of the reference run on it unmodified.  The fixture holds the
(user, item, rating) triples of those 20 000 rows (so the tests never read /root/reference), every 5th online prediction plus
the six the survey tabulates, the cumulative MSE, and rotation-invariant summaries of the final state.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ml100k.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _import_reference   # noqa: E402

N, NU, NI, ETA, M, STRIDE = 20000, 943, 1682, 0.005, 5, 5
KAT_IDX = [0, 1, 10, 100, 1000, 19999]


def triples():
    raw = np.loadtxt(os.path.join(REF, "dataset", "ml-100k", "ua.base"), dtype=np.int64)
    order = np.argsort(raw[:, 3].astype(np.float32), kind="stable")[:N]
    return raw[order, 0] - 1, raw[order, 1] - 1, raw[order, 2]


def onehot(u, it):
    X = np.zeros((len(u), NU + NI + 1))
    X[np.arange(len(u)), u] = 1
    X[np.arange(len(u)), NU + it] = 1
    X[:, -1] = 1
    return X


def probe(d, cols=4):
    return np.random.RandomState(2626).standard_normal((d, cols))


def main():
    _import_reference()
    from models.models_online.FM_FTRL import FM_FTRL
    from models.models_online.SFTRL_CCFM import SFTRL_CCFM
    from models.models_online.SFTRL_Vanila import SFTRL_Vanila
    T = torch.DoubleTensor
    u, it, r = triples()
    X, y = onehot(u, it), r.astype(np.float64)
    out = {"user": u.astype(np.int16), "item": it.astype(np.int16), "rating": r.astype(np.int8),
           "meta": np.array([N, NU, NI, ETA, M, STRIDE], np.float64), "kat_idx": np.array(KAT_IDX)}
    Z = probe(X.shape[1])
    with contextlib.redirect_stdout(io.StringIO()):
        for tag, cls in (("ccfm", SFTRL_CCFM), ("vanila", SFTRL_Vanila)):
            mdl = cls(T(X), T(y), "reg", ETA, M)
            pred, _, _ = mdl.online_learning()
            pred = np.asarray([float(p) for p in pred])
            out[tag + "_pred"] = pred[::STRIDE].copy()
            out[tag + "_kat"] = pred[KAT_IDX].copy()
            out[tag + "_mse"] = np.array([np.mean((pred - y) ** 2)])
            out[tag + "_rc"] = np.array([mdl.row_count_p, mdl.row_count_n])
            for key, BT in (("BTP", mdl.BT_P.numpy()), ("BTN", mdl.BT_N.numpy())):
                out[f"{tag}_{key}_sv"] = np.linalg.svd(BT, compute_uv=False)
                out[f"{tag}_{key}_probe"] = BT @ (BT.T @ Z[:BT.shape[0]])       # (BT BT^T) Z: rotation-invariant
            if tag == "vanila":
                out[tag + "_w"] = mdl.w.numpy().copy()
        torch.manual_seed(0)
        mdl = FM_FTRL(T(X), T(y), "reg", ETA, M)
        torch.manual_seed(0)
        mdl._init_parameter()
        # torch.randn draws fp32 and the reference casts to fp64 (FM_FTRL.py:42-43): fp32 storage is exact
        out["ftrl_w1_init"] = mdl.w1.numpy().astype(np.float32)
        out["ftrl_W2_init"] = mdl.W2.numpy().astype(np.float32)
        assert np.array_equal(out["ftrl_W2_init"].astype(np.float64), mdl.W2.numpy())
        torch.manual_seed(0)
        pred, _, _ = mdl.online_learning()
        pred = np.asarray([float(p) for p in pred])
        out["ftrl_pred"] = pred[::STRIDE].copy()
        out["ftrl_kat"] = pred[KAT_IDX].copy()
        out["ftrl_mse"] = np.array([np.mean((pred - y) ** 2)])
        out["ftrl_w1"] = mdl.w1.numpy().copy()
        out["ftrl_W2_probe"] = mdl.W2.numpy() @ Z[:-1]
    for tag in ("ccfm", "vanila", "ftrl"):
        print(tag, "mse %.10f" % out[tag + "_mse"][0], "kat", np.array2string(out[tag + "_kat"], precision=8))
    np.savez_compressed(os.path.join(HERE, "ml100k_kat.npz"), **out)


if __name__ == "__main__":
    main()
