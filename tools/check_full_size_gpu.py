"""Full-size GPU checks against live-reference fixtures that were written AFTER the round's GPU budget was spent.

The three CUDA classical learners:
  * BASELINE.json configs[1] (59 535 x 8, m = 40, 'cls'): tests/golden/classical_full.npz
    (CPU twin: test_oracle_matches_reference_at_full_size_cfg2);
  * SURVEY.md section 4's known answers on the bundled ml-100k rows (20 000 x 2626 one-hot, 'reg', m = 5):
    tests/golden/ml100k_kat.npz (CPU twin: test_oracle_reproduces_the_ml100k_known_answers).

The five CUDA deep classes through the experiment scripts' flow at its real length (100 pre-training steps on a 2 500-sample
batch, run_experiment over a 2 500-sample stream): tests/golden/online_full.npz (CPU twin: tests/test_online_full.py).

NOT part of `pytest -m gpu` yet: none of this has run on a B200.  Promote each part to a test once
`gpurun -- python tools/check_full_size_gpu.py` has printed CFG2_FULL_OK, ML100K_KAT_OK, ONLINE_FULL_OK and SGD_TRAJ_OK
(the last one: 1 000-step plain-SGD trajectories, tests/golden/traj_sgd.npz).
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import GOLDEN, auc, rmse                      # noqa: E402
from golden.make_golden_classical import codrna          # noqa: E402


def main():
    import torch
    import fm_for_online_recommendation_b200 as pkg
    g = dict(np.load(GOLDEN + "/classical_full.npz"))
    N, seed, eta, m, _, stride = (float(v) for v in g["meta"])
    N, m, stride = int(N), int(m), int(stride)
    X, y = codrna(N, int(seed))
    T = torch.DoubleTensor

    def stream(pred):
        pred = np.asarray(pred, np.float64)
        return [round(auc(pred, y), 4), round(float(np.mean(np.sign(pred) == np.sign(y))), 4), round(rmse(pred, y), 4)]

    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(7)
        mdl = pkg.FM_FTRL(T(X), T(y), "cls", eta, m)
        p, real, secs = mdl.online_learning()
    p = np.asarray(p, np.float64).reshape(-1)
    print("FM_FTRL     %.3f s  sign mismatches (every %dth): %d" % (secs, stride, int((p[::stride] != g["ftrl_pred"]).sum())))
    assert stream(p) == [round(float(v), 4) for v in g["ftrl_metrics"]]
    np.testing.assert_allclose(mdl.w1.cpu().numpy(), g["ftrl_w1"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(mdl.W2.cpu().numpy(), g["ftrl_W2"], rtol=1e-9, atol=1e-12)
    for tag, cls in (("ccfm", pkg.SFTRL_CCFM), ("vanila", pkg.SFTRL_Vanila)):
        with contextlib.redirect_stdout(io.StringIO()):
            mdl = cls(T(X), T(y), "cls", eta, m)
            p, _, secs = mdl.online_learning()
        p = np.asarray(p, np.float64).reshape(-1)
        print("SFTRL_%-6s %.3f s  sign mismatches: %d  rows %s / %s" % (tag, secs, int((p[::stride] != g[tag + "_pred"]).sum()),
                                                                        [mdl.row_count_p, mdl.row_count_n], g[tag + "_rc"].tolist()))
        assert stream(p) == [round(float(v), 4) for v in g[tag + "_metrics"]]
        assert [mdl.row_count_p, mdl.row_count_n] == g[tag + "_rc"].tolist()
        for key, mine in (("BTP", mdl.BT_P), ("BTN", mdl.BT_N)):
            ref = g[f"{tag}_{key}"]
            mine = mine.cpu().numpy()
            np.testing.assert_allclose(mine @ mine.T, ref @ ref.T, rtol=1e-8, atol=1e-10)
    np.random.seed(5)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        mdl = pkg.RRF_Online(T(X), T(y), "cls", num_sampled_spectral=10)
        assert np.array_equal(mdl.gamma.cpu().numpy(), g["rrf_gamma0"]) and np.array_equal(mdl.eps.cpu().numpy(), g["rrf_eps"])
        p, _, secs = mdl.online_learning()
    p = np.asarray(p, np.float64).reshape(-1)
    print("RRF_Online  %.3f s  sign mismatches: %d" % (secs, int((p[::stride] != g["rrf_pred"]).sum())))
    assert len(p) == int(g["rrf_n"][0]) and np.array_equal(p[::stride], g["rrf_pred"])
    np.testing.assert_allclose(mdl.w.cpu().numpy(), g["rrf_w"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(mdl.gamma.cpu().numpy(), g["rrf_gamma"], rtol=1e-8, atol=1e-11)
    print("CFG2_FULL_OK")


def ml100k():
    import torch
    import fm_for_online_recommendation_b200 as pkg
    from test_classical import SURVEY_KAT, ml100k_case
    g, X, y, eta, m, stride, Z = ml100k_case()
    kat = g["kat_idx"]
    T = torch.DoubleTensor
    for tag, cls in (("ccfm", pkg.SFTRL_CCFM), ("vanila", pkg.SFTRL_Vanila), ("ftrl", pkg.FM_FTRL)):
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(0)
            mdl = cls(T(X), T(y), "reg", eta, m)
            p, _, secs = mdl.online_learning()
        p = np.asarray(p, np.float64).reshape(-1)
        mse, preds = SURVEY_KAT[tag]
        print("%-6s %.3f s  mse %.10f (survey %.10f)  max |pred - reference| %.2e" %
              (tag, secs, np.mean((p - y) ** 2), mse, np.abs(p[::stride] - g[tag + "_pred"]).max()))
        assert abs(np.mean((p - y) ** 2) - mse) < 1e-8
        np.testing.assert_allclose(p[kat], preds, rtol=0, atol=1e-8)
        np.testing.assert_allclose(p[::stride], g[tag + "_pred"], rtol=1e-9, atol=1e-9)
        if tag != "ftrl":
            assert [mdl.row_count_p, mdl.row_count_n] == g[tag + "_rc"].tolist()
            for key, BT in (("BTP", mdl.BT_P), ("BTN", mdl.BT_N)):
                BT = BT.cpu().numpy()
                np.testing.assert_allclose(BT @ (BT.T @ Z[:BT.shape[0]]), g[f"{tag}_{key}_probe"], rtol=1e-8, atol=1e-9)
    print("ML100K_KAT_OK")


def online_full():
    import torch
    import fm_for_online_recommendation_b200 as pkg
    from test_online_full import CFG, G
    from traj_common import batch, init_tables
    pre, lr, L, H = G["meta"]
    w1, V = init_tables(CFG)
    pXi, pXv, pY = batch(CFG, 0)
    oXi, oXv, oY = batch(CFG, 1)
    for kind in ("FMAdam", "DeepFMAdam", "NFMAdam", "DeepFMOnn", "NFMOnn"):
        kw = dict(embedding_size=10, n=float(lr))
        if kind != "FMAdam":
            kw.update(num_hidden_layers=int(L), neuron_per_hidden_layer=int(H))
        m = getattr(pkg, kind)(CFG["sizes"], **kw)
        with torch.no_grad():
            t = torch.zeros_like(m._table)
            t[:, :10] = torch.from_numpy(V)
            t[:, 10] = torch.from_numpy(w1)
            m._table.copy_(t)
            m.bias.copy_(torch.from_numpy(G[kind + "_init_bias"]).reshape(m.bias.shape))
            if kind != "FMAdam":
                m._mlp.copy_(torch.from_numpy(G[kind + "_init_mlp"]))
            if kind + "_init_alpha" in G:
                m.alpha.copy_(torch.from_numpy(G[kind + "_init_alpha"]))
        losses = np.asarray([float(m.update_embedding(pXi, pXv, pY).cpu()) for _ in range(int(pre))], np.float32)
        secs, acc, roc, conf = m.run_experiment(oXi, oXv, [int(v) for v in oY])
        want = dict(zip(("tp", "fp", "tn", "fn"), G[kind + "_conf"].tolist()))
        print("%-10s losses bit-equal: %s   confusion %s (reference %s)   %.3f s" %
              (kind, np.array_equal(losses, G[kind + "_pre_loss"]), conf, want, secs))
        assert np.array_equal(losses, G[kind + "_pre_loss"]) and conf == want
        assert [acc, roc["tpr"], roc["fpr"]] == G[kind + "_acc_roc"].tolist()
        rows = G["rows"]
        tab = m._table.cpu().numpy()
        dv = np.abs(tab[rows, :10] - G[kind + "_V"]) / np.maximum(np.abs(G[kind + "_V"]), 1e-3)
        assert dv.max() <= 1e-5, dv.max()
    print("ONLINE_FULL_OK")


def sgd_trajectories():
    """tests/golden/traj_sgd.npz (CPU twin: test_oracle_reproduces_sgd_trajectories_bit_for_bit)"""
    import torch
    import fm_for_online_recommendation_b200 as pkg
    from _util import synth
    from golden.make_trajectory_sgd import CASES, CKPT, cfg_of
    from traj_common import digest, init_tables
    g = dict(np.load(GOLDEN + "/traj_sgd.npz"))
    for name, (kind, sizes, B, zipf, (L, H), lr, steps) in CASES.items():
        kw = dict(embedding_size=10, n=lr, update_mode=1)
        if L:
            kw.update(num_hidden_layers=L, neuron_per_hidden_layer=H)
        m = getattr(pkg, kind)(sizes, **kw)
        w1, V = init_tables(cfg_of(name))
        with torch.no_grad():
            t = torch.zeros_like(m._table)
            t[:, :10] = torch.from_numpy(V)
            t[:, 10] = torch.from_numpy(w1)
            m._table.copy_(t)
            m.bias.copy_(torch.from_numpy(g[name + "_init_bias"]).reshape(m.bias.shape))
            if L:
                m._mlp.copy_(torch.from_numpy(g[name + "_init_mlp"]))
        losses = []
        for s in range(steps):
            Xi, Xv, Y = synth(sizes, B, 7000 + s, real_xv=name.endswith("_xv"), zipf=zipf)
            losses.append(float(m.update_embedding(Xi, Xv, Y).cpu()))
            if (s + 1) in CKPT:
                tab = m._table.cpu().numpy()
                d = digest(np.ascontiguousarray(tab[:, :10]), np.ascontiguousarray(tab[:, 10]), m.bias.detach().cpu().numpy().reshape(1))
                assert d == str(g["%s_s%d_digest" % (name, s + 1)]), (name, s + 1)
        assert np.array_equal(np.asarray(losses, np.float32), g[name + "_losses"]), name
        print("%-15s %d SGD steps bit-identical with the live reference + torch.optim.SGD" % (name, steps))
    print("SGD_TRAJ_OK")


if __name__ == "__main__":
    main()
    ml100k()
    online_full()
    sgd_trajectories()
