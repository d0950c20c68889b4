"""Golden fixtures for RRF_Online, generated from the REFERENCE itself (models/models_online/RRF_Online.py; build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_rrf.py
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
from make_golden_classical import codrna, onehot  # noqa: E402


def main():
    _import_reference()
    from models.models_online.RRF_Online import RRF_Online
    T = torch.DoubleTensor
    cases = {"codrna_cls": (codrna(2000, 10), "cls", 10), "onehot_reg": (onehot(1200, 30, 40, 11), "reg", 6),
             "codrna_reg": ((codrna(1500, 12)[0], onehot(1500, 3, 3, 13)[1]), "reg", 10)}
    out = {}
    for name, ((X, y), task, D) in cases.items():
        np.random.seed(5)
        torch.manual_seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            m = RRF_Online(T(X), T(y), task, num_sampled_spectral=D)
            out[name + "_gamma0"] = m.gamma.numpy().copy()
            out[name + "_w0"] = m.w.numpy().copy()
            out[name + "_eps"] = m.eps.numpy().copy()
            pred, real, _ = m.online_learning()
        out[name + "_X"], out[name + "_y"] = X, y
        out[name + "_meta"] = np.array([0 if task == "reg" else 1, D], np.int64)
        out[name + "_pred"] = np.asarray([float(p) for p in pred])
        out[name + "_gamma"] = m.gamma.numpy().copy()
        out[name + "_w"] = m.w.numpy().copy()
        print("done", name, len(pred), float(np.mean((out[name + "_pred"] - y[:len(pred)]) ** 2)))
    np.savez_compressed(os.path.join(HERE, "rrf.npz"), **out)


if __name__ == "__main__":
    main()
