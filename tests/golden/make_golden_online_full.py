"""The experiment scripts' flow at its REAL length, from the live reference (build container only): the five deep classes
as main_experiment.py:51-84 constructs them (k = 10, 5 hidden layers of 10, n = 1e-4, raw N(0,1) embeddings),
100 pre-training `update_embedding` steps on one 2 500-sample batch (:89-96; the script does 1 000) and then
`run_experiment` over a 2 500-sample stream (:123-128; predict -> fit per example).

Inputs are regenerated from seeds (tests/traj_common.py); the embedding tables are seeded from numpy so that they need not be
stored.  The fixture holds the tower / alpha / bias initial values, the pre-training losses, the confusion counts, accuracy
and ROC point of the stream, 512 sampled table rows + the dense parameters of the final state and held-out scores.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_online_full.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import _import_reference, flat_params   # noqa: E402
from traj_common import EVAL_STEP, batch, init_tables, sample_rows   # noqa: E402

CFG = dict(sizes=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], B=2500, seed=31, scale=None, kw=dict(embedding_size=10))
PRE, LR, L, H = 100, 1e-4, 5, 10
KINDS = ["FMAdam", "DeepFMAdam", "NFMAdam", "DeepFMOnn", "NFMOnn"]


def set_tables(model, w1, V):
    with torch.no_grad():
        o = 0
        for e1, e2 in zip(model.first_order_embeddings, model.second_order_embeddings):
            n = e1.weight.shape[0]
            e1.weight.copy_(torch.from_numpy(w1[o:o + n, None]))
            e2.weight.copy_(torch.from_numpy(V[o:o + n]))
            o += n


def main():
    _import_reference()
    import importlib
    torch.set_num_threads(1)
    mods = {"FMAdam": "fm_adam", "DeepFMAdam": "deepfm_adam", "NFMAdam": "nfm_adam", "DeepFMOnn": "deepfm_onn",
            "NFMOnn": "nfm_onn"}
    w1, V = init_tables(CFG)
    rows = sample_rows(V.shape[0], 512)
    pXi, pXv, pY = batch(CFG, 0)
    oXi, oXv, oY = batch(CFG, 1)
    eXi, eXv, _ = batch(CFG, EVAL_STEP)
    out = {"rows": rows, "meta": np.array([PRE, LR, L, H], np.float64)}
    for kind in KINDS:
        cls = getattr(importlib.import_module("models.models_online_deep." + mods[kind]), kind)
        kw = dict(embedding_size=10, n=LR)
        if kind != "FMAdam":
            kw.update(num_hidden_layers=L, neuron_per_hidden_layer=H)
        torch.manual_seed(5)
        m = cls(CFG["sizes"], use_cuda=False, **kw)
        set_tables(m, w1, V)
        p = flat_params(m)
        for key in ("mlp", "bias", "alpha"):
            if key in p:
                out[f"{kind}_init_{key}"] = p[key]
        out[kind + "_pre_loss"] = np.asarray(
            [float(m.update_embedding(pXi.tolist(), pXv.tolist(), pY.tolist()).detach()) for _ in range(PRE)], np.float32)
        _, acc, roc, conf = m.run_experiment(oXi.tolist(), oXv.tolist(), [int(v) for v in oY])
        out[kind + "_conf"] = np.asarray([conf["tp"], conf["fp"], conf["tn"], conf["fn"]], np.int64)
        out[kind + "_acc_roc"] = np.asarray([acc, roc["tpr"], roc["fpr"]], np.float64)
        q = flat_params(m)
        out[kind + "_V"], out[kind + "_w1"] = q["V"][rows], q["w1"][rows]
        for key in ("mlp", "bias", "alpha"):
            if key in q:
                out[f"{kind}_{key}"] = q[key]
        with torch.no_grad():
            f = m.forward(eXi.tolist(), eXv.tolist())
            out[kind + "_eval_z"] = (f[0] if isinstance(f, tuple) else f).numpy().astype(np.float32)
        print(kind, conf, acc, flush=True)
    np.savez_compressed(os.path.join(HERE, "online_full.npz"), **out)


if __name__ == "__main__":
    main()
