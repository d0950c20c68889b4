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
    # RRF_Online over the same 59 535-sample stream
    p, st = oc.rrf_online(X, y, "cls", g["rrf_gamma0"], g["rrf_w0"], g["rrf_eps"])
    assert len(p) == int(g["rrf_n"][0])
    assert np.array_equal(p[::stride], g["rrf_pred"])
    np.testing.assert_allclose(st["w"], g["rrf_w"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st["gamma"], g["rrf_gamma"], rtol=1e-9, atol=1e-12)


def ml100k_case():
    """first 20 000 time-ordered ratings of the reference's bundled ml-100k `ua.base` as the one-hot user | item | bias matrix
    of utils/data_manager.py:18-48 (d = 2626), from the committed triples"""
    g = dict(np.load(GOLDEN + "/ml100k_kat.npz"))
    N, NU, NI, eta, m, stride = (float(v) for v in g["meta"])
    N, NU, NI, m, stride = int(N), int(NU), int(NI), int(m), int(stride)
    u, it = g["user"].astype(np.int64), g["item"].astype(np.int64)
    X = np.zeros((N, NU + NI + 1))
    X[np.arange(N), u] = 1
    X[np.arange(N), NU + it] = 1
    X[:, -1] = 1
    Z = np.random.RandomState(2626).standard_normal((X.shape[1], 4))
    return g, X, g["rating"].astype(np.float64), eta, m, stride, Z


# known answers on the bundled data (reg, eta = 0.005, m = 5), "expect agreement to ~1e-8"
SURVEY_KAT = {
    "ccfm": (1.0218898321, [0, 0.16, 0.31228865, 3.64095492, 2.95804486, 3.83643576]),
    "vanila": (1.0447981956, [0, 0.12, 0.34185643, 3.49008605, 3.08330353, 3.95040156]),
    "ftrl": (1.0822492538, [18.44102295, -0.2867719, 0.20292217, 3.17376672, 3.11692213, 3.99502804]),
}


def test_oracle_reproduces_the_ml100k_known_answers():
    """the survey's known-answer table AND the fixture generated from the live reference on the same rows
    (tests/golden/make_golden_ml100k.py): cumulative MSE, tabulated predictions, every 5th prediction, final state"""
    g, X, y, eta, m, stride, Z = ml100k_case()
    kat = g["kat_idx"]

    def check_stream(tag, p):
        mse, preds = SURVEY_KAT[tag]
        assert abs(np.mean((p - y) ** 2) - mse) < 1e-8 and abs(np.mean((p - y) ** 2) - g[tag + "_mse"][0]) < 1e-11
        np.testing.assert_allclose(p[kat], preds, rtol=0, atol=1e-8)
        np.testing.assert_allclose(p[kat], g[tag + "_kat"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(p[::stride], g[tag + "_pred"], rtol=1e-10, atol=1e-10)
        assert round(rmse(p, y), 4) == round(float(np.sqrt(g[tag + "_mse"][0])), 4)

    for tag, van in (("ccfm", False), ("vanila", True)):
        p, st = oc.sftrl(X, y, "reg", eta, m, vanila=van)
        check_stream(tag, p)
        assert [st["row_count_p"], st["row_count_n"]] == g[tag + "_rc"].tolist()
        for key, BT in (("BTP", st["BT_P"]), ("BTN", st["BT_N"])):
            np.testing.assert_allclose(np.linalg.svd(BT, compute_uv=False), g[f"{tag}_{key}_sv"], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(BT @ (BT.T @ Z[:BT.shape[0]]), g[f"{tag}_{key}_probe"], rtol=1e-9, atol=1e-10)
        if van:
            np.testing.assert_allclose(st["w"].reshape(-1), g["vanila_w"].reshape(-1), rtol=1e-10, atol=1e-13)
    p, st = oc.fm_ftrl(X, y, "reg", eta, m, g["ftrl_w1_init"].astype(np.float64), g["ftrl_W2_init"].astype(np.float64))
    check_stream("ftrl", p)
    np.testing.assert_allclose(st["w1"], g["ftrl_w1"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(st["W2"] @ Z[:-1], g["ftrl_W2_probe"], rtol=1e-11, atol=1e-14)


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
