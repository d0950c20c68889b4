"""Metrics on the device against the reference's Python loops (utils/metric_manager.py:7-29, restated
here line by line: the reference cannot be imported on the GPU box) and against the exact rank-statistic AUC."""
import numpy as np
import pytest
import torch

from _util import auc as auc_ref, synth

pytestmark = pytest.mark.gpu


def ref_regression_metric(pred, real):          # utils/metric_manager.py:7-15
    val = 0.0
    metric = [np.inf]
    for idx, (p, r) in enumerate(zip(pred, real)):
        val += (p - r) ** 2
        metric.append((1 / (idx + 1)) * val)
    return np.asarray(metric).reshape([-1, 1])


def ref_classfication_metric(pred, real):       # utils/metric_manager.py:18-29
    val = 0.0
    metric, metric_acc = [], []
    for idx, (p, r) in enumerate(zip(pred, real)):
        val += 1 if p == r else 0
        metric.append(1 / (idx + 1) * np.log(1.0 + np.exp(-p * r)))
        metric_acc.append(1 / (idx + 1) * val)
    return np.asarray(metric).reshape([-1, 1]), np.asarray(metric_acc).reshape([-1, 1])


@pytest.mark.parametrize("n", [1, 31, 33, 5000])
def test_running_curves_equal_the_reference_loops(n):
    from fm_for_online_recommendation_b200 import metrics
    rng = np.random.RandomState(n)
    pred = rng.standard_normal(n) * 3
    real = rng.randint(1, 6, size=n).astype(np.float64)
    got = metrics.regression_metric(pred, real)
    want = ref_regression_metric(pred, real)
    assert got.shape == want.shape and np.isinf(got[0, 0])
    assert np.array_equal(got[1:], want[1:])            # same fp64 operations in the same order: bit-identical
    pc = np.where(rng.uniform(size=n) < 0.5, 1.0, -1.0)
    rc = np.where(rng.uniform(size=n) < 0.5, 1.0, -1.0)
    m, a = metrics.classfication_metric(pc, rc)
    wm, wa = ref_classfication_metric(pc, rc)
    assert np.array_equal(a, wa)
    np.testing.assert_allclose(m, wm, rtol=1e-13)        # log/exp of the device maths library vs numpy's


def test_exact_auc_and_confusion_counts():
    from fm_for_online_recommendation_b200 import metrics
    rng = np.random.RandomState(0)
    for n in (2, 257, 20000):
        s = np.round(rng.standard_normal(n), 1).astype(np.float32)     # rounded: plenty of ties
        y = (rng.uniform(size=n) < 0.3).astype(np.float32)
        if y.min() == y.max():
            y[0], y[-1] = 0, 1
        assert abs(metrics.auc(s, y) - auc_ref(s, y)) < 1e-12
        pred = s > 0
        c = metrics.confusion(pred, y)
        assert c == {"tp": int((pred & (y == 1)).sum()), "fp": int((pred & (y == 0)).sum()),
                     "tn": int((~pred & (y == 0)).sum()), "fn": int((~pred & (y == 1)).sum())}


def test_run_experiment_bookkeeping_comes_from_the_device_counts():
    """run_experiment returns (seconds, accuracy[-1], roc[-1], confusion) of fm_adam.py:101-119 computed from the kernel's
    own counters; they must equal the reference loop evaluated on the returned predictions."""
    import fm_for_online_recommendation_b200 as pkg
    sizes = [7, 5, 11, 3, 13, 4]
    torch.manual_seed(3)
    m = pkg.FMAdam(sizes, embedding_size=6, n=0.01)
    with torch.no_grad():
        m._table.mul_(0.2)
    Xi, Xv, Y = synth(sizes, 2311, 9)
    _, acc, roc, conf = m.run_experiment(Xi.tolist(), Xv.tolist(), [int(v) for v in Y])
    preds = m._last_online_preds
    cm = {"tp": 0, "fp": 0, "tn": 0, "fn": 0}
    for p, yv in zip(preds, Y):                      # fm_adam.py:101-111
        if p == yv:
            cm["tp" if yv == 1 else "tn"] += 1
        else:
            cm["fn" if yv == 1 else "fp"] += 1
    assert conf == cm
    assert acc == (cm["tp"] + cm["tn"]) / len(Y) * 100
    assert roc == {"tpr": cm["tp"] / (cm["tp"] + cm["fn"] + 1e-16), "fpr": cm["fp"] / (cm["fp"] + cm["tn"] + 1e-16)}
    # predict_proba: the scores whose threshold predict() returns
    Xi2, Xv2, _ = synth(sizes, 300, 10)
    p = m.predict_proba(Xi2, Xv2).cpu().numpy()
    assert np.array_equal(p > 0.5, m.predict(Xi2, Xv2))
