"""single-pass vs L-pass hedge fit at cfg4's shape: difference of the tower weights after 1, 2, 3 fits"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from _util import synth, rel_err
from test_gpu_fm import CRITEO
import fm_for_online_recommendation_b200 as pkg
B = 8192
res = {}
for single in ("1", "0"):
    os.environ["FMB_HEDGE_SINGLE"] = single
    torch.manual_seed(5)
    m = pkg.DeepFMOnn(CRITEO, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-2, batch_size=B)
    with torch.no_grad():
        m._table[:, :11].mul_(0.05)
    w0 = m._mlp.cpu().numpy().copy()
    for step in range(3):
        Xi, Xv, Y = synth(CRITEO, B, 70 + step)
        m.fit(m.encode(Xi, Xv, Y), None, None)
        res[(single, step)] = (m._mlp.cpu().numpy().copy(), m.alpha.detach().cpu().numpy().copy())
for step in range(3):
    a, b = res[("1", step)][0], res[("0", step)][0]
    upd = np.abs(b - w0).max()
    print("step", step, "rel_err", rel_err(a, b), "max abs diff", np.abs(a - b).max(), "max |update so far|", upd,
          "diff/update", np.abs(a - b).max() / upd, "alpha", res[("1", step)][1], res[("0", step)][1])
