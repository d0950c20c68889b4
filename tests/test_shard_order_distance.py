"""How far the multi-GPU summation orders sit from the reference order (CPU, oracle only).

The CUDA ranks are bit-exact with the oracle's restatement of THEIR order (tests/test_sharded*.py); these tests pin the
relation between that order and the reference order that tools/shard_order_distance.py measures at full size
(profiles/r2_shard_order_distance_*.json):
  * G = 1 and two-field models: the orders coincide bit for bit (a two-term sum commutes);
  * otherwise one step moves a coordinate by at most the sign step's 2 * lr, and the held-out metrics stay together
    while individual weights separate (the sign step amplifies one-ulp differences, like it does inside the reference
    itself between two hosts).
"""
import numpy as np
import pytest

from _util import auc, rmse, synth

K, LR = 6, 1e-3
TOL_METRIC = 2e-3     # held-out AUC / RMSE after 60 steps: measured 1e-4 .. 4e-4 at these sizes


def _pair(sizes, order, G, B):
    from oracle.deep import OracleDeep
    out = []
    for o in ("reference", order):
        orc = OracleDeep("FMAdam", sizes, K, lr=LR, seed=3)
        orc.V *= np.float32(0.2)
        if o == "owner":
            orc.set_shard_order(G)
        elif o == "rank_partial":
            orc.set_rank_partial_order(B // G)
        out.append(orc)
    return out


def _steps(ref, other, sizes, B, steps):
    gaps = []
    for s in range(steps):
        Xi, _, Y = synth(sizes, B, 900 + s, zipf=(s % 3 == 2))
        Xv = np.ones(Xi.shape, np.float32)
        before = np.array_equal(ref.V, other.V) and np.array_equal(ref.w1, other.w1) and np.array_equal(ref.bias, other.bias)
        ref.update_embedding(Xi, Xv, Y)
        other.update_embedding(Xi, Xv, Y)
        if before:      # from identical state one step differs by at most a flipped sign step
            gaps.append(max(float(np.abs(ref.V - other.V).max()), float(np.abs(ref.w1 - other.w1).max())))
    return gaps


@pytest.mark.parametrize("order,G", [("owner", 1), ("owner", 2), ("owner", 8), ("rank_partial", 1)])
def test_orders_that_coincide_with_the_reference_order(order, G):
    sizes, B = ([9, 40, 7, 3, 100, 23, 2, 64] if G == 1 else [943, 1682]), 64
    ref, other = _pair(sizes, order, G, B)
    _steps(ref, other, sizes, B, 40)
    assert np.array_equal(ref.V, other.V) and np.array_equal(ref.w1, other.w1) and np.array_equal(ref.bias, other.bias)


@pytest.mark.parametrize("order,G", [("owner", 2), ("owner", 8), ("rank_partial", 2), ("rank_partial", 8)])
def test_reordered_sums_stay_within_the_sign_step_and_keep_the_metrics(order, G):
    sizes, B = [9, 40, 7, 3, 100, 23, 2, 64, 11, 5, 300, 17], 128
    ref, other = _pair(sizes, order, G, B)
    gaps = _steps(ref, other, sizes, B, 60)
    assert gaps, "the first step starts from identical tables"
    assert max(gaps) <= 2 * LR * (1 + 1e-6), max(gaps)
    eXi, _, eY = synth(sizes, 4096, 7)
    eXv = np.ones(eXi.shape, np.float32)
    zr, zo = ref.forward_fm(eXi, eXv), other.forward_fm(eXi, eXv)
    pr, po = (1.0 / (1.0 + np.exp(-z.astype(np.float64))) for z in (zr, zo))
    assert abs(auc(zr, eY) - auc(zo, eY)) <= TOL_METRIC
    assert abs(rmse(pr, eY) - rmse(po, eY)) <= TOL_METRIC
