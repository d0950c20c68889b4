"""BASELINE.json configs[1] at its FULL size on the GPU (59 535 x 8, m = 40, 'cls'): the three CUDA learners against the
live-reference fixture tests/golden/classical_full.npz (CPU twin: test_oracle_matches_reference_at_full_size_cfg2).

NOT part of `pytest -m gpu` yet: written after the round's GPU budget was spent, so it has not run on a B200.  Promote it to a
test once `gpurun -- python tools/check_cfg2_full.py` has printed CFG2_FULL_OK.
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
    print("CFG2_FULL_OK")


if __name__ == "__main__":
    main()
