"""The experiment scripts' flow at its real length (main_experiment.py:51-128): five deep classes, 100 pre-training steps on
a 2 500-sample batch, then `run_experiment` over a 2 500-sample stream (predict -> fit per example) -- the oracle against
the LIVE reference's record (tests/golden/make_golden_online_full.py).  Bars: pre-training losses bit for bit, confusion
counts / accuracy / ROC point identical, table rows within 1e-5 (north_star), tower weights and tower-model scores within
1e-5 on >= 99.8 % / 99.9 % of the entries (worst measured values next to the tolerances below: MKL's sgemm order is not
mirrored), AUC / RMSE of the held-out scores equal to 4 decimals.  The CUDA persistent kernel is compared bit for bit with this oracle in
tests/test_gpu_deep.py (run_experiment, 300-sample streams)."""
import numpy as np
import pytest

from _util import GOLDEN, auc, rel_err, rmse
from traj_common import EVAL_STEP, batch, init_tables

CFG = dict(sizes=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], B=2500, seed=31, scale=None, kw=dict(embedding_size=10))
TOL = 1e-5
TOL_TOWER = 1e-3   # worst tower weight after 100 + 2 500 steps, relative to max(|w|, 1e-3): measured 2.4e-4 (|diff| 2.4e-7) on
                   # ONE of NFMAdam's 550 weights, 1.1e-5 on one of DeepFMAdam's; every other weight within 1e-5
G = dict(np.load(GOLDEN + "/online_full.npz"))


@pytest.mark.parametrize("kind", ["FMAdam", "DeepFMAdam", "NFMAdam", "DeepFMOnn", "NFMOnn"])
def test_oracle_reproduces_the_scripts_flow_at_full_length(kind):
    from oracle.deep import OracleDeep
    pre, lr, L, H = G["meta"]
    pre, L, H = int(pre), (0 if kind == "FMAdam" else int(L)), (0 if kind == "FMAdam" else int(H))
    orc = OracleDeep(kind, CFG["sizes"], 10, L, H, lr=float(lr), **(dict(batch_size=1) if "Onn" in kind else {}))
    orc.w1[:], orc.V[:] = init_tables(CFG)
    orc.bias[:] = G[kind + "_init_bias"]
    if L:
        orc.mlp[:] = G[kind + "_init_mlp"]
    if kind + "_init_alpha" in G:
        orc.alpha[:] = G[kind + "_init_alpha"]
    pXi, pXv, pY = batch(CFG, 0)
    losses = np.asarray([orc.update_embedding(pXi, pXv, pY) for _ in range(pre)], np.float32)
    assert np.array_equal(losses, G[kind + "_pre_loss"])
    oXi, oXv, oY = batch(CFG, 1)
    conf, preds = orc.run_experiment(oXi, oXv, oY)
    want = dict(zip(("tp", "fp", "tn", "fn"), G[kind + "_conf"].tolist()))
    assert conf == want
    acc = (conf["tp"] + conf["tn"]) / len(oY) * 100
    tpr = conf["tp"] / (conf["tp"] + conf["fn"] + 1e-16)
    fpr = conf["fp"] / (conf["fp"] + conf["tn"] + 1e-16)
    assert [acc, tpr, fpr] == G[kind + "_acc_roc"].tolist()          # fm_adam.py:108-116
    rows = G["rows"]
    assert rel_err(orc.V[rows], G[kind + "_V"]) <= TOL and rel_err(orc.w1[rows], G[kind + "_w1"]) <= TOL
    assert rel_err(orc.bias, G[kind + "_bias"]) <= TOL
    if L:   # MKL's sgemm order is not mirrored: a sign step now and then lands one ulp apart on a near-zero tower weight
        a, b = orc.mlp.astype(np.float64), G[kind + "_mlp"].astype(np.float64)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        print(kind, "tower weights: max rel", rel.max(), "beyond 1e-5:", int((rel > TOL).sum()), "of", rel.size)
        assert (rel > TOL).mean() <= 0.002 and rel.max() <= TOL_TOWER
    if kind + "_alpha" in G:
        assert rel_err(orc.alpha, G[kind + "_alpha"]) <= TOL
    eXi, eXv, eY = batch(CFG, EVAL_STEP)
    f = orc.forward(eXi, eXv)
    z = f[0] if isinstance(f, tuple) else f
    ref = G[kind + "_eval_z"]
    rz = np.abs(z.astype(np.float64) - ref) / np.maximum(np.abs(ref.astype(np.float64)), 1e-3)
    if L:   # tower models: scores near zero carry the tower's sgemm-order difference (measured worst 1.07e-5, NFMAdam)
        assert (rz > TOL).mean() <= 0.001 and rz.max() <= 1e-4
    else:
        assert rz.max() <= TOL
    if np.isfinite(ref).all() and len(np.unique(eY)) > 1:
        assert round(auc(z, eY), 4) == round(auc(ref, eY), 4)
        sg = lambda v: 1.0 / (1.0 + np.exp(-np.asarray(v, np.float64)))
        assert round(rmse(sg(z), eY), 4) == round(rmse(sg(ref), eY), 4)
