"""Shared helpers for the test-suite (golden replay, synthetic generators)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_DEEP = {
    "fm_cfg1": ("FMAdam", dict(k=10)),
    "fm_small_scaled": ("FMAdam", dict(k=10)),
    "deepfm_small": ("DeepFMAdam", dict(k=10, L=3, H=16)),
    "deepfm_raw": ("DeepFMAdam", dict(k=10, L=2, H=8)),
    "nfm_k64": ("NFMAdam", dict(k=64, L=1, H=64)),
    "deepfm_onn": ("DeepFMOnn", dict(k=10, L=5, H=10, batch_size=1)),
    "nfm_onn": ("NFMOnn", dict(k=10, L=5, H=10, batch_size=1)),
    "nfm_onn_b8": ("NFMOnn", dict(k=10, L=3, H=10, batch_size=8)),
}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3))) if a.size else 0.0


def synth(feature_sizes, n, seed, real_xv=False, zipf=False):
    rng = np.random.RandomState(seed)
    if zipf:
        Xi = np.stack([np.minimum(rng.zipf(1.3, size=n) - 1, fs - 1) for fs in feature_sizes], 1)
    else:
        Xi = np.stack([rng.randint(0, fs, size=n) for fs in feature_sizes], 1)
    Xv = rng.uniform(0.25, 2.0, size=Xi.shape).astype(np.float32) if real_xv else np.ones(Xi.shape, np.float32)
    Y = (rng.uniform(size=n) < 0.3).astype(np.float32)
    return Xi.astype(np.int64), Xv, Y


def auc(scores, labels):
    """exact ROC AUC (rank statistic with average ranks for ties), labels in {0,1} or {-1,+1}."""
    s = np.asarray(scores, np.float64).reshape(-1)
    y = np.asarray(labels).reshape(-1) > 0
    order = np.argsort(s, kind="stable")
    ranks = np.empty(len(s), np.float64)
    ss = s[order]
    i = 0
    while i < len(ss):
        j = i
        while j + 1 < len(ss) and ss[j + 1] == ss[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    npos, nneg = int(y.sum()), int((~y).sum())
    if npos == 0 or nneg == 0:
        return float("nan")
    return float((ranks[y].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))


def rmse(pred, target):
    d = np.asarray(pred, np.float64).reshape(-1) - np.asarray(target, np.float64).reshape(-1)
    return float(np.sqrt(np.mean(d * d)))
