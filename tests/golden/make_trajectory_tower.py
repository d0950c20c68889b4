"""1 000-step `fit` trajectories of the tower models from the LIVE reference (build container only).

tests/golden/make_trajectory.py covers the FM-only steps (bit-exact over 10 000 steps); here the whole model trains:
DeepFMAdam.fit / NFMAdam.fit (deepfm_adam.py:106-117, nfm_adam.py:105-116) and the hedge step DeepFMOnn.fit
(deepfm_onn.py:109-154) at B = 256 with a 3 x 32 tower on Frappe-shaped fields.  The fixture holds the tower / bias / alpha
initial values and, at steps 300 and 1 000, 512 sampled table rows, the dense parameters and the held-out scores.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_trajectory_tower.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import _import_reference, flat_params            # noqa: E402
from make_golden_online_full import set_tables                    # noqa: E402
from traj_common import EVAL_STEP, batch, init_tables, sample_rows   # noqa: E402

CFG = dict(sizes=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], B=256, seed=41, scale=0.2, kw=dict(embedding_size=10))
L, H, LR, STEPS, CKPT = 3, 32, 1e-3, 1000, (300, 1000)
MODS = {"DeepFMAdam": "deepfm_adam", "NFMAdam": "nfm_adam", "DeepFMOnn": "deepfm_onn"}


def main():
    _import_reference()
    torch.set_num_threads(1)
    w1, V = init_tables(CFG)
    rows = sample_rows(V.shape[0], 512)
    eXi, eXv, _ = batch(CFG, EVAL_STEP)
    out = {"rows": rows, "meta": np.array([L, H, LR, STEPS], np.float64), "ckpt": np.array(CKPT)}
    for kind, mod in MODS.items():
        cls = getattr(importlib.import_module("models.models_online_deep." + mod), kind)
        kw = dict(embedding_size=10, n=LR, num_hidden_layers=L, neuron_per_hidden_layer=H)
        if "Onn" in kind:
            kw["batch_size"] = CFG["B"]
        torch.manual_seed(5)
        m = cls(CFG["sizes"], use_cuda=False, **kw)
        set_tables(m, w1, V)
        p = flat_params(m)
        for key in ("mlp", "bias", "alpha"):
            if key in p:
                out[f"{kind}_init_{key}"] = p[key]
        for s in range(STEPS):
            Xi, Xv, Y = batch(CFG, s)
            m.fit(Xi.tolist(), Xv.tolist(), Y.tolist())
            if (s + 1) in CKPT:
                q = flat_params(m)
                tag = f"{kind}_s{s + 1}_"
                out[tag + "V"], out[tag + "w1"] = q["V"][rows], q["w1"][rows]
                for key in ("mlp", "bias", "alpha"):
                    if key in q:
                        out[tag + key] = q[key]
                with torch.no_grad():
                    f = m.forward(eXi.tolist(), eXv.tolist())
                    out[tag + "eval_z"] = (f[0] if isinstance(f, tuple) else f).numpy().astype(np.float32)
        print("done", kind, flush=True)
    np.savez_compressed(os.path.join(HERE, "traj_tower_fit.npz"), **out)


if __name__ == "__main__":
    main()
