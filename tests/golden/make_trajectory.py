"""Generate the long-horizon trajectory fixtures from the REFERENCE itself (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_trajectory.py [name ...]

Runs the reference's own classes (models/models_online_deep/*.py, use_cuda=False, one thread) for the step
counts of tests/traj_common.TRAJ -- 10 000 update_embedding steps at BASELINE.json configs[0]'s shape, 1 000 at
configs[2]'s and configs[3]'s -- and stores, per checkpoint: every loss so far, the parameters (in full when small,
otherwise a SHA-256 over the packed tables plus a sample of rows), and AUC / RMSE of the scores on a held-out
batch.  tests/test_trajectory.py replays the same batches through the oracle (CPU) and the CUDA path (GPU) and
demands bit-identical losses and parameters at every checkpoint.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True
from make_golden import _import_reference, flat_params  # noqa: E402
from traj_common import EVAL_STEP, TRAJ, batch, digest, init_tables, sample_rows, sizes_of  # noqa: E402
from _util import auc, rmse  # noqa: E402


def run(name):
    cfg = TRAJ[name]
    import importlib
    mod = {"FMAdam": "fm_adam", "DeepFMAdam": "deepfm_adam", "NFMAdam": "nfm_adam"}[cfg["kind"]]
    cls = getattr(importlib.import_module("models.models_online_deep." + mod), cfg["kind"])
    torch.manual_seed(cfg["seed"])
    torch.set_num_threads(1)   # ATen's chunking of element-wise kernels depends on the thread count above 32768 elements
    model = cls(sizes_of(cfg), n=cfg["lr"], use_cuda=False, **cfg["kw"])
    w1, V = init_tables(cfg)
    with torch.no_grad():
        o = 0
        for f, fs in enumerate(sizes_of(cfg)):
            model.first_order_embeddings[f].weight.copy_(torch.from_numpy(w1[o:o + fs, None]))
            model.second_order_embeddings[f].weight.copy_(torch.from_numpy(V[o:o + fs]))
            o += fs
    out = {"init_bias": flat_params(model)["bias"]}
    R = w1.shape[0]
    rows = sample_rows(R)
    big = R > 3000
    eXi, eXv, eY = batch(cfg, EVAL_STEP)
    losses = []
    t0 = time.time()
    torch.set_num_threads(int(os.environ.get('TRAJ_THREADS', 8 if R > 100000 else 1)))   # big: dense Adam over 11 M parameters
    for s in range(cfg["steps"]):
        Xi, Xv, Y = batch(cfg, s)
        if cfg["method"] == "update_embedding":
            losses.append(float(model.update_embedding(Xi, Xv, Y).detach()))
        else:
            model.fit(Xi, Xv, Y)
        if (s + 1) in cfg["ckpt"]:
            p = flat_params(model)
            tag = "s%d_" % (s + 1)
            out[tag + "digest"] = np.array(digest(p["V"], p["w1"], p["bias"]))
            if big:
                out[tag + "rows"] = rows
                out[tag + "V"] = p["V"][rows]
                out[tag + "w1"] = p["w1"][rows]
            else:
                out[tag + "V"] = p["V"]
                out[tag + "w1"] = p["w1"]
            out[tag + "bias"] = p["bias"]
            with torch.no_grad():
                z = (model.forward_fm(eXi, eXv) if hasattr(model, "forward_fm") else model.forward(eXi, eXv)).numpy()
            out[tag + "eval_z"] = z
            out[tag + "auc"] = np.float64(auc(z, eY))
            out[tag + "rmse"] = np.float64(rmse(1.0 / (1.0 + np.exp(-z.astype(np.float64))), eY))
            print(name, "step", s + 1, "loss", losses[-1] if losses else None, "auc %.6f rmse %.6f" %
                  (out[tag + "auc"], out[tag + "rmse"]), "%.0fs" % (time.time() - t0), flush=True)
    out["losses"] = np.asarray(losses, np.float32)
    np.savez_compressed(os.path.join(HERE, "traj_" + name + ".npz"), **out)


if __name__ == "__main__":
    _import_reference()
    for n in (sys.argv[1:] or list(TRAJ)):
        run(n)
