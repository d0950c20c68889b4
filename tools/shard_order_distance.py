"""Distance between the multi-GPU summation orders and the reference order.

The row-sharded step sums the FM forward terms owner by owner (csrc/sharded.cu, csrc/shard3.cu: `set_shard_order(G)`),
and csrc/shard2.cu sums duplicate rows rank-partial first (`set_rank_partial_order(B / G)`).  Both orders are restated in
the oracle and the CUDA ranks are bit-exact with THAT; this script measures how far those orders drift from the
reference order (the single-GPU step, bit-identical with the live reference over 10 000 steps, tests/test_trajectory.py)
on the same seeded trajectories, next to a noise floor (`ulp_G1`: the reference order itself, started one ulp away in 1 % of
the coordinates).  CPU only (the oracle); nothing of the product runs here.

    python tools/shard_order_distance.py [--traj cfg1_ue] [--steps 3000] [--out profiles/r2_shard_order_distance.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _util import auc, rmse                                          # noqa: E402
from traj_common import EVAL_STEP, TRAJ, batch, init_tables, sizes_of   # noqa: E402


def make(cfg, order, G):
    from oracle.deep import OracleDeep
    kw = cfg["kw"]
    orc = OracleDeep(cfg["kind"], sizes_of(cfg), kw["embedding_size"], kw.get("num_hidden_layers", 0),
                     kw.get("neuron_per_hidden_layer", 0), lr=cfg["lr"])
    orc.w1[:], orc.V[:] = init_tables(cfg)
    orc.bias[:] = np.float32(0.99)
    if order == "owner":
        orc.set_shard_order(G)
    elif order == "rank_partial":
        orc.set_rank_partial_order(cfg["B"] // G)
    elif order == "ulp":
        # noise floor: the REFERENCE order started one ulp away in 1 % of the embedding coordinates -- what any rounding-level
        # difference (a different thread count in the reference's own torch.sum, say) grows into under the sign step
        rng = np.random.RandomState(5)
        flat = orc.V.reshape(-1)
        pick = rng.choice(flat.size, size=max(1, flat.size // 100), replace=False)
        flat[pick] = np.nextafter(flat[pick], np.float32(np.inf) * np.where(rng.uniform(size=pick.size) < 0.5, -1, 1).astype(np.float32))
    return orc


def compare(ref, other, eXi, eXv, eY):
    tab_r = np.concatenate([ref.V.reshape(-1), ref.w1.reshape(-1), ref.bias.reshape(-1)])
    tab_o = np.concatenate([other.V.reshape(-1), other.w1.reshape(-1), other.bias.reshape(-1)])
    d = np.abs(tab_r.astype(np.float64) - tab_o.astype(np.float64))
    rel = d / np.maximum(np.abs(tab_r.astype(np.float64)), 1e-3)
    zr, zo = ref.forward_fm(eXi, eXv), other.forward_fm(eXi, eXv)
    pr, po = (1.0 / (1.0 + np.exp(-z.astype(np.float64))) for z in (zr, zo))
    return {
        "bit_equal_frac": float((tab_r == tab_o).mean()),
        "max_abs": float(d.max()), "max_rel": float(rel.max()),
        "frac_rel_gt_1e-5": float((rel > 1e-5).mean()),
        "eval_logit_max_abs": float(np.abs(zr.astype(np.float64) - zo.astype(np.float64)).max()),
        "auc_ref": round(auc(zr, eY), 6), "auc": round(auc(zo, eY), 6),
        "rmse_ref": round(rmse(pr, eY), 6), "rmse": round(rmse(po, eY), 6),
        "auc_rmse_equal_4dp": bool(round(auc(zr, eY), 4) == round(auc(zo, eY), 4)
                                   and round(rmse(pr, eY), 4) == round(rmse(po, eY), 4)),
    }


def run(name, steps, ckpts, variants):
    cfg = TRAJ[name]
    assert cfg["method"] == "update_embedding"
    ref = make(cfg, "reference", 1)
    others = {"%s_G%d" % v: make(cfg, *v) for v in variants}
    eXi, eXv, eY = batch(cfg, EVAL_STEP)
    out = {"trajectory": name, "B": cfg["B"], "lr": cfg["lr"], "fields": len(cfg["sizes"]), "checkpoints": {}}
    loss_gap = {n: 0.0 for n in others}
    for s in range(steps):
        Xi, Xv, Y = batch(cfg, s)
        lr_ = ref.update_embedding(Xi, Xv, Y)
        for n, o in others.items():
            lo = o.update_embedding(Xi, Xv, Y)
            loss_gap[n] = max(loss_gap[n], abs(float(lo) - float(lr_)) / max(abs(float(lr_)), 1e-3))
        if (s + 1) in ckpts:
            out["checkpoints"][s + 1] = {n: dict(compare(ref, o, eXi, eXv, eY), loss_max_rel_so_far=loss_gap[n])
                                         for n, o in others.items()}
            print(json.dumps({s + 1: out["checkpoints"][s + 1]}), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", default="cfg1_ue")
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    os.environ.setdefault("ORC_THREADS", "8")
    ck = [c for c in (1, 10, 100, 1000, 3000, 10000) if c <= a.steps]
    res = run(a.traj, a.steps, ck, [("owner", 2), ("owner", 8), ("rank_partial", 2), ("rank_partial", 8), ("ulp", 1)])
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
