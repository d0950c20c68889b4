"""Classical online learners (A9-A11): the numpy oracle is pinned against the reference's outputs
(CPU test), the CUDA persistent kernels are compared with the oracle and the golden (GPU test).
fp64 tolerance: 1e-9 relative on regression scores, identical +-1 decisions for 'cls'
(dgemv/LAPACK summation orders are not mirrored; sketches are compared through BT BT^T, which is
invariant to the sign/rotation freedom of singular vectors -- SURVEY.md section 7)."""
import contextlib
import io

import numpy as np
import pytest

from _util import GOLDEN, auc, rmse
from oracle import classical as oc

CASES = ["codrna_cls_m40", "onehot_reg_m5", "onehot_cls_m3"]
G = dict(np.load(GOLDEN + "/classical.npz"))


def case(name):
    eta, m, t = G[name + "_meta"]
    return G[name + "_X"], G[name + "_y"], ("cls" if t else "reg"), float(eta), int(m)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    X, y, task, eta, m = case(name)
    p, st = oc.fm_ftrl(X, y, task, eta, m, G[name + "_ftrl_w1_init"], G[name + "_ftrl_W2_init"])
    np.testing.assert_allclose(p, G[name + "_ftrl_pred"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(st["W2"], G[name + "_ftrl_W2"], rtol=1e-12, atol=1e-14)
    for tag, van in (("ccfm", False), ("vanila", True)):
        p, st = oc.sftrl(X, y, task, eta, m, vanila=van)
        np.testing.assert_allclose(p, G[f"{name}_{tag}_pred"], rtol=1e-10, atol=1e-12)
        assert [st["row_count_p"], st["row_count_n"]] == G[f"{name}_{tag}_rc"].tolist()
        for key, mine in (("BTP", st["BT_P"]), ("BTN", st["BT_N"])):
            ref = G[f"{name}_{tag}_{key}"]
            np.testing.assert_allclose(mine @ mine.T, ref @ ref.T, rtol=1e-9, atol=1e-12)


def test_oracle_matches_reference_at_full_size_cfg2():
    """BASELINE.json configs[1] at its full size (59 535 cod-rna-shaped samples x 8 features, m = 40, 'cls'): fixture from the
    live reference (tests/golden/make_golden_classical_full.py: every 8th online prediction, final state, stream metrics)."""
    from golden.make_golden_classical import codrna
    g = dict(np.load(GOLDEN + "/classical_full.npz"))
    N, seed, eta, m, _, stride = (float(v) for v in g["meta"])
    N, m, stride = int(N), int(m), int(stride)
    assert N == 59535 and m == 40
    X, y = codrna(N, int(seed))

    def stream(pred):
        pred = np.asarray(pred, np.float64)
        return [auc(pred, y), float(np.mean(np.sign(pred) == np.sign(y))), rmse(pred, y)]

    p, st = oc.fm_ftrl(X, y, "cls", eta, m, g["ftrl_w1_init"], g["ftrl_W2_init"])
    assert np.array_equal(p[::stride], g["ftrl_pred"])                   # 'cls' predictions are signs
    assert [round(v, 4) for v in stream(p)] == [round(float(v), 4) for v in g["ftrl_metrics"]]
    np.testing.assert_allclose(st["w1"], g["ftrl_w1"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st["W2"], g["ftrl_W2"], rtol=1e-12, atol=1e-14)
    for tag, van in (("ccfm", False), ("vanila", True)):
        p, st = oc.sftrl(X, y, "cls", eta, m, vanila=van)
        assert np.array_equal(p[::stride], g[tag + "_pred"])
        assert [round(v, 4) for v in stream(p)] == [round(float(v), 4) for v in g[tag + "_metrics"]]
        assert [st["row_count_p"], st["row_count_n"]] == g[tag + "_rc"].tolist()
        for key, mine in (("BTP", st["BT_P"]), ("BTN", st["BT_N"])):
            ref = g[f"{tag}_{key}"]
            np.testing.assert_allclose(mine @ mine.T, ref @ ref.T, rtol=1e-9, atol=1e-11)
        if van:
            np.testing.assert_allclose(st["w"].reshape(-1), g["vanila_w"].reshape(-1), rtol=1e-9, atol=1e-13)


def metrics_match(pred, ref, y, task):
    """north_star: AUC and RMSE identical to 4 decimal places (online prediction streams)."""
    if task == "reg":
        assert round(rmse(pred, y), 4) == round(rmse(ref, y), 4)
    else:
        assert round(auc(pred, y), 4) == round(auc(ref, y), 4)
        assert np.mean(np.sign(pred) == np.sign(y)) == np.mean(np.sign(ref) == np.sign(y))


@pytest.mark.parametrize("name", CASES)
def test_oracle_auc_rmse_identical_to_4_decimals(name):
    X, y, task, eta, m = case(name)
    p, _ = oc.fm_ftrl(X, y, task, eta, m, G[name + "_ftrl_w1_init"], G[name + "_ftrl_W2_init"])
    metrics_match(p, G[name + "_ftrl_pred"], y, task)
    for tag, van in (("ccfm", False), ("vanila", True)):
        p, _ = oc.sftrl(X, y, task, eta, m, vanila=van)
        metrics_match(p, G[f"{name}_{tag}_pred"], y, task)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_oracle_and_reference(name):
    import torch
    import fm_for_online_recommendation_b200 as pkg
    X, y, task, eta, m = case(name)
    T = torch.DoubleTensor
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(7)
        mdl = pkg.FM_FTRL(T(X), T(y), task, eta, m)
        p, real, _ = mdl.online_learning()
    np.testing.assert_allclose(p, G[name + "_ftrl_pred"], rtol=1e-9, atol=1e-9)
    metrics_match(np.asarray(p), G[name + "_ftrl_pred"], y, task)
    np.testing.assert_allclose(mdl.W2.cpu().numpy(), G[name + "_ftrl_W2"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(mdl.w1.cpu().numpy(), G[name + "_ftrl_w1"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(real, y)
    for tag, cls, van in (("ccfm", pkg.SFTRL_CCFM, False), ("vanila", pkg.SFTRL_Vanila, True)):
        with contextlib.redirect_stdout(io.StringIO()):
            mdl = cls(T(X), T(y), task, eta, m)
            p, _, _ = mdl.online_learning()
        po, st = oc.sftrl(X, y, task, eta, m, vanila=van)
        np.testing.assert_allclose(p, po, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(p, G[f"{name}_{tag}_pred"], rtol=1e-9, atol=1e-9)
        metrics_match(np.asarray(p), G[f"{name}_{tag}_pred"], y, task)
        assert [mdl.row_count_p, mdl.row_count_n] == G[f"{name}_{tag}_rc"].tolist()
        for key, mine in (("BTP", mdl.BT_P), ("BTN", mdl.BT_N)):
            ref = G[f"{name}_{tag}_{key}"]
            mine = mine.cpu().numpy()
            np.testing.assert_allclose(mine @ mine.T, ref @ ref.T, rtol=1e-8, atol=1e-11)
        if van:
            np.testing.assert_allclose(mdl.w.cpu().numpy().reshape(-1), G[f"{name}_{tag}_w"].reshape(-1), rtol=1e-9,
                                       atol=1e-12)


@pytest.mark.gpu
def test_nan_raises_value_error():
    import torch
    import fm_for_online_recommendation_b200 as pkg
    X = np.random.RandomState(0).uniform(-1, 1, (50, 6))
    X[10, 2] = np.nan
    y = np.ones(50)
    with contextlib.redirect_stdout(io.StringIO()):
        mdl = pkg.SFTRL_CCFM(torch.DoubleTensor(X), torch.DoubleTensor(y), "reg", 0.01, 3)
        with pytest.raises(ValueError, match="Nan contained"):
            mdl.online_learning()
