"""Full-size fixture for BASELINE.json configs[1] (cod-rna shape: 59 535 samples x 8 features, m = 40, 'cls'), generated
from the REFERENCE itself (models/models_online/{FM_FTRL,SFTRL_CCFM,SFTRL_Vanila}.py; build container only).

The inputs are regenerated from the seed, so the fixture holds only the reference's outputs: every 8th online prediction,
the final state, and AUC / accuracy / RMSE of the whole prediction stream.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_classical_full.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import _import_reference        # noqa: E402
from make_golden_classical import codrna         # noqa: E402

N, SEED, TASK, ETA, M, STRIDE = 59535, 100, "cls", 0.005, 40, 8


def stream_metrics(pred, y):
    from _util import auc, rmse
    pred = np.asarray(pred, np.float64)
    return np.array([auc(pred, y), float(np.mean(np.sign(pred) == np.sign(y))), rmse(pred, y)], np.float64)


def main():
    _import_reference()
    from models.models_online.FM_FTRL import FM_FTRL
    from models.models_online.SFTRL_CCFM import SFTRL_CCFM
    from models.models_online.SFTRL_Vanila import SFTRL_Vanila
    T = torch.DoubleTensor
    X, y = codrna(N, SEED)
    out = {"meta": np.array([N, SEED, ETA, M, 1, STRIDE], np.float64)}
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.time()
        torch.manual_seed(7)
        mdl = FM_FTRL(T(X), T(y), TASK, ETA, M)
        torch.manual_seed(7)
        mdl._init_parameter()
        out["ftrl_w1_init"] = mdl.w1.numpy().copy()
        out["ftrl_W2_init"] = mdl.W2.numpy().copy()
        torch.manual_seed(7)
        pred, _, _ = mdl.online_learning()
        pred = np.asarray([float(p) for p in pred])
        out["ftrl_pred"] = pred[::STRIDE].copy()
        out["ftrl_metrics"] = stream_metrics(pred, y)
        out["ftrl_w1"] = mdl.w1.numpy().copy()
        out["ftrl_W2"] = mdl.W2.numpy().copy()
        out["ftrl_seconds"] = np.array([time.time() - t0])
        for tag, cls in (("ccfm", SFTRL_CCFM), ("vanila", SFTRL_Vanila)):
            t0 = time.time()
            mdl = cls(T(X), T(y), TASK, ETA, M)
            pred, _, _ = mdl.online_learning()
            pred = np.asarray([float(p) for p in pred])
            out[f"{tag}_pred"] = pred[::STRIDE].copy()
            out[f"{tag}_metrics"] = stream_metrics(pred, y)
            out[f"{tag}_BTP"] = mdl.BT_P.numpy().copy()
            out[f"{tag}_BTN"] = mdl.BT_N.numpy().copy()
            out[f"{tag}_rc"] = np.array([mdl.row_count_p, mdl.row_count_n])
            if tag == "vanila":
                out[f"{tag}_w"] = mdl.w.numpy().copy()
            out[f"{tag}_seconds"] = np.array([time.time() - t0])
        # RRF_Online on the same stream; the constructor draws gamma / w / eps from the CPU RNGs
        from models.models_online.RRF_Online import RRF_Online
        t0 = time.time()
        np.random.seed(5)
        torch.manual_seed(5)
        mdl = RRF_Online(T(X), T(y), TASK, num_sampled_spectral=10)
        out["rrf_gamma0"], out["rrf_w0"], out["rrf_eps"] = mdl.gamma.numpy().copy(), mdl.w.numpy().copy(), mdl.eps.numpy().copy()
        pred, _, _ = mdl.online_learning()
        pred = np.asarray([float(p) for p in pred])
        out["rrf_n"] = np.array([len(pred)])
        out["rrf_pred"] = pred[::STRIDE].copy()
        out["rrf_metrics"] = stream_metrics(pred, y[:len(pred)])
        out["rrf_w"], out["rrf_gamma"] = mdl.w.numpy().copy(), mdl.gamma.numpy().copy()
        out["rrf_seconds"] = np.array([time.time() - t0])
    for k in ("ftrl", "ccfm", "vanila", "rrf"):
        print(k, "reference seconds on this host:", float(out[k + "_seconds"][0]), "metrics", out[k + "_metrics"])
    np.savez_compressed(os.path.join(HERE, "classical_full.npz"), **out)


if __name__ == "__main__":
    main()
