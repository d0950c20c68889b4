"""Plain-SGD trajectories from the LIVE reference's forward pass + torch autograd + torch.optim.SGD (build container only).

BASELINE.json configs[0] is "FM k = 10 offline SGD" on ml-100k-shaped ids; the reference's classes only ever build a fresh
Adam (fm_adam.py:60), whose first step is a SIGN step that hides the gradient's magnitude.  Under SGD every bit of the
gradient reaches the weights, so these trajectories pin the VALUES of the sparse gradient (BCE backward, the duplicate-row
summation order of torch's embedding_dense_backward) and update mode 1 (`fma(g, -lr, p)`), not just their signs.
Harness: fm_adam.py:56-69 with `torch.optim.SGD(model.parameters(), lr)` in place of the Adam of line 60.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_trajectory_sgd.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from _util import synth                                           # noqa: E402
from make_golden import _import_reference, flat_params            # noqa: E402
from make_golden_online_full import set_tables                    # noqa: E402
from traj_common import digest, init_tables, sample_rows                       # noqa: E402

FRAPPE = [957, 4082, 7, 7, 2, 3, 2, 9, 80, 233]
# name -> (class, feature_sizes, B, zipf ids, (L, H), lr, steps)
CASES = {
    "cfg1": ("FMAdam", [943, 1682], 256, False, (0, 0), 0.01, 1000),
    "frappe_zipf": ("FMAdam", FRAPPE, 256, True, (0, 0), 0.01, 1000),       # many duplicate rows per batch
    "deepfm_fm_part": ("DeepFMAdam", FRAPPE, 250, True, (2, 8), 0.01, 1000),  # forward_fm of a tower class; B % 32 != 0
    # NFM's update_embedding puts the loss on sigmoid(z) (nfm_adam.py:99): the second loss variant, backward through sigmoid
    "nfm_loss_of_sigmoid": ("NFMAdam", FRAPPE, 250, True, (1, 16), 0.05, 1000),
    # real-valued Xv (data_preprocess.py:113: the svm reader's values): a name ending in _xv draws Xv from U(0.25, 2)
    "frappe_zipf_xv": ("FMAdam", FRAPPE, 256, True, (0, 0), 0.01, 1000),
}
CKPT = (1, 100, 1000)


def cfg_of(name):
    kind, sizes, B, zipf, (L, H), lr, steps = CASES[name]
    return dict(sizes=sizes, B=B, seed=51, scale=0.2, kw=dict(embedding_size=10))


def main():
    _import_reference()
    from models.models_online_deep.deepfm_adam import DeepFMAdam
    from models.models_online_deep.fm_adam import FMAdam
    from models.models_online_deep.nfm_adam import NFMAdam
    torch.set_num_threads(1)
    out = {}
    for name, (kind, sizes, B, zipf, (L, H), lr, steps) in CASES.items():
        kw = dict(embedding_size=10, n=lr)
        if L:
            kw.update(num_hidden_layers=L, neuron_per_hidden_layer=H)
        torch.manual_seed(5)
        m = {"FMAdam": FMAdam, "DeepFMAdam": DeepFMAdam, "NFMAdam": NFMAdam}[kind](sizes, use_cuda=False, **kw)
        set_tables(m, *init_tables(cfg_of(name)))
        p = flat_params(m)
        out[name + "_init_bias"], out[name + "_init_mlp"] = p["bias"], p["mlp"]
        opt = torch.optim.SGD(m.parameters(), lr=lr)
        losses = []
        for s in range(steps):
            Xi, Xv, Y = synth(sizes, B, 7000 + s, real_xv=name.endswith("_xv"), zipf=zipf)
            opt.zero_grad()
            z = m.forward_fm(Xi.tolist(), Xv.tolist()) if hasattr(m, "forward_fm") else m.forward(Xi.tolist(), Xv.tolist())
            loss = F.binary_cross_entropy_with_logits(torch.sigmoid(z) if kind == "NFMAdam" else z, torch.from_numpy(Y))
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
            if (s + 1) in CKPT:
                q = flat_params(m)
                out[f"{name}_s{s + 1}_digest"] = np.array(digest(q["V"], q["w1"], q["bias"]))
                if s + 1 == steps:      # the digest covers the whole table; a sample of rows helps when it does not match
                    rows = sample_rows(q["V"].shape[0], 256)
                    out[name + "_rows"], out[name + "_V"], out[name + "_w1"], out[name + "_bias"] = \
                        rows, q["V"][rows], q["w1"][rows], q["bias"]
        out[name + "_losses"] = np.asarray(losses, np.float32)
        print("done", name, losses[0], losses[-1], flush=True)
    np.savez_compressed(os.path.join(HERE, "traj_sgd.npz"), **out)


if __name__ == "__main__":
    main()
