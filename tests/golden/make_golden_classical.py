"""Golden fixtures for the classical online learners, generated from the REFERENCE itself
(models/models_online/{FM_FTRL,SFTRL_CCFM,SFTRL_Vanila}.py; run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_classical.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402


def codrna(n, seed):
    """cod-rna-shaped stream: 8 features in U(-1,1), y = +1 w.p. 1/3 else -1."""
    rng = np.random.RandomState(seed)
    X = rng.uniform(-1, 1, size=(n, 8))
    y = np.where(rng.uniform(size=n) < 1 / 3, 1.0, -1.0)
    return X, y


def onehot(n, nu, ni, seed):
    """ml-100k-shaped stream: one-hot user | item | bias column, integer ratings 1..5."""
    rng = np.random.RandomState(seed)
    X = np.zeros((n, nu + ni + 1))
    u = rng.randint(0, nu, n)
    it = rng.randint(0, ni, n)
    X[np.arange(n), u] = 1
    X[np.arange(n), nu + it] = 1
    X[:, -1] = 1
    y = rng.randint(1, 6, n).astype(np.float64)
    return X, y


def main():
    _import_reference()
    from models.models_online.FM_FTRL import FM_FTRL
    from models.models_online.SFTRL_CCFM import SFTRL_CCFM
    from models.models_online.SFTRL_Vanila import SFTRL_Vanila
    T = torch.DoubleTensor
    cases = {
        "codrna_cls_m40": (codrna(3000, 0), "cls", 0.005, 40),
        "onehot_reg_m5": (onehot(1500, 30, 40, 1), "reg", 0.005, 5),
        "onehot_cls_m3": ((onehot(800, 12, 9, 2)[0], codrna(800, 3)[1]), "cls", 0.05, 3),
    }
    out = {}
    for name, ((X, y), task, eta, m) in cases.items():
        out[name + "_X"] = X
        out[name + "_y"] = y
        out[name + "_meta"] = np.array([eta, m, 0 if task == "reg" else 1], np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(7)
            mdl = FM_FTRL(T(X), T(y), task, eta, m)
            torch.manual_seed(7)
            mdl._init_parameter()
            out[name + "_ftrl_w1_init"] = mdl.w1.numpy().copy()
            out[name + "_ftrl_W2_init"] = mdl.W2.numpy().copy()
            torch.manual_seed(7)
            pred, _, _ = mdl.online_learning()
            out[name + "_ftrl_pred"] = np.asarray([float(p) for p in pred])
            out[name + "_ftrl_w1"] = mdl.w1.numpy().copy()
            out[name + "_ftrl_W2"] = mdl.W2.numpy().copy()
            for tag, cls in (("ccfm", SFTRL_CCFM), ("vanila", SFTRL_Vanila)):
                mdl = cls(T(X), T(y), task, eta, m)
                pred, _, _ = mdl.online_learning()
                out[f"{name}_{tag}_pred"] = np.asarray([float(p) for p in pred])
                out[f"{name}_{tag}_BTP"] = mdl.BT_P.numpy().copy()
                out[f"{name}_{tag}_BTN"] = mdl.BT_N.numpy().copy()
                out[f"{name}_{tag}_rc"] = np.array([mdl.row_count_p, mdl.row_count_n])
                if tag == "vanila":
                    out[f"{name}_{tag}_w"] = mdl.w.numpy().copy()
        print("done", name)
    np.savez_compressed(os.path.join(HERE, "classical.npz"), **out)


if __name__ == "__main__":
    main()
