"""How far the oracle's tower `fit` drifts from the LIVE reference over long horizons (build container only: imports
/root/reference).  The FM-only steps are bit-identical with the reference over 10 000 steps (tests/test_trajectory.py); the
tower's products are not (MKL's sgemm summation order is not mirrored), and the sign step amplifies every last-bit difference.
This script puts numbers on that for DeepFMAdam.fit / NFMAdam.fit / DeepFMOnn.fit at the reference scripts' tower shapes.

    PYTHONDONTWRITEBYTECODE=1 python tools/tower_fit_drift.py [--steps 3000] [--out profiles/r2_tower_fit_drift.json]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from _util import auc, rmse                                   # noqa: E402
from make_golden import _import_reference, flat_params        # noqa: E402
from traj_common import EVAL_STEP, batch, init_tables         # noqa: E402

MODS = {"DeepFMAdam": "deepfm_adam", "NFMAdam": "nfm_adam", "DeepFMOnn": "deepfm_onn", "NFMOnn": "nfm_onn"}


def run(kind, B, L, H, lr, scale, steps, ckpts):
    import torch
    from oracle.deep import OracleDeep
    from make_golden_online_full import set_tables
    cfg = dict(sizes=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], B=B, seed=41, scale=scale, kw=dict(embedding_size=10))
    cls = getattr(importlib.import_module("models.models_online_deep." + MODS[kind]), kind)
    torch.manual_seed(5)
    torch.set_num_threads(1)
    kw = dict(embedding_size=10, n=lr, num_hidden_layers=L, neuron_per_hidden_layer=H)
    if "Onn" in kind:
        kw["batch_size"] = B
    ref = cls(cfg["sizes"], use_cuda=False, **kw)
    w1, V = init_tables(cfg)
    set_tables(ref, w1, V)
    orc = OracleDeep(kind, cfg["sizes"], 10, L, H, lr=lr, **(dict(batch_size=B) if "Onn" in kind else {}))
    p = flat_params(ref)
    orc.w1[:], orc.V[:], orc.bias[:], orc.mlp[:] = p["w1"], p["V"], p["bias"], p["mlp"]
    if "alpha" in p:
        orc.alpha[:] = p["alpha"]
    eXi, eXv, eY = batch(cfg, EVAL_STEP)
    out = {}
    for s in range(steps):
        Xi, Xv, Y = batch(cfg, s)
        ref.fit(Xi.tolist(), Xv.tolist(), Y.tolist())
        orc.fit(Xi, Xv, Y)
        if (s + 1) in ckpts:
            q = flat_params(ref)
            rel = lambda a, b: np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b.astype(np.float64)), 1e-3)
            rV, rM = rel(orc.V, q["V"]), rel(orc.mlp, q["mlp"])
            with torch.no_grad():
                f = ref.forward(eXi.tolist(), eXv.tolist())
                zr = (f[0] if isinstance(f, tuple) else f).numpy()
            f = orc.forward(eXi, eXv)
            zo = f[0] if isinstance(f, tuple) else f
            sg = lambda v: 1.0 / (1.0 + np.exp(-np.asarray(v, np.float64)))
            out[s + 1] = {"table_bit_equal": float((orc.V == q["V"]).mean()), "table_gt_1e-5": float((rV > 1e-5).mean()),
                          "table_max_rel": float(rV.max()), "tower_gt_1e-5": float((rM > 1e-5).mean()),
                          "tower_max_rel": float(rM.max()), "score_max_abs": float(np.abs(zr.astype(np.float64) - zo).max()),
                          "auc_ref": round(auc(zr, eY), 6), "auc": round(auc(zo, eY), 6),
                          "rmse_ref": round(rmse(sg(zr), eY), 6), "rmse": round(rmse(sg(zo), eY), 6)}
            print(kind, s + 1, json.dumps(out[s + 1]), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    _import_reference()
    ck = [c for c in (1, 10, 100, 300, 1000, 3000, 10000) if c <= a.steps]
    res = {"config": "Frappe-shaped fields (5 382 rows), k = 10, B = 256, tower 3 x 32, lr = 1e-3, tables N(0,1) * 0.2",
           "runs": {}}
    for kind in ("DeepFMAdam", "NFMAdam", "DeepFMOnn"):
        res["runs"][kind] = run(kind, 256, 3, 32, 1e-3, 0.2, a.steps, ck)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
