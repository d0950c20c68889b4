"""RRF_Online (models/models_online/RRF_Online.py:70-187): the numpy oracle is pinned on the reference's
outputs (CPU), the persistent CUDA kernel is compared with the oracle and the golden (GPU).  fp64, tolerance 1e-9
relative on regression scores and on the final state, identical +-1 decisions for 'cls' (the order of the d-term dot
products and cos/sin/exp implementations are not mirrored)."""
import contextlib
import io

import numpy as np
import pytest

from _util import GOLDEN
from oracle import classical as oc

CASES = ["codrna_cls", "onehot_reg", "codrna_reg"]
G = dict(np.load(GOLDEN + "/rrf.npz"))


def case(name):
    t, D = G[name + "_meta"]
    return G[name + "_X"], G[name + "_y"], ("cls" if t else "reg"), int(D)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    X, y, task, D = case(name)
    p, st = oc.rrf_online(X, y, task, G[name + "_gamma0"], G[name + "_w0"], G[name + "_eps"])
    assert len(p) == len(G[name + "_pred"])
    np.testing.assert_allclose(p, G[name + "_pred"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(st["w"], G[name + "_w"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st["gamma"], G[name + "_gamma"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_oracle_and_reference(name):
    import torch
    import fm_for_online_recommendation_b200 as pkg
    X, y, task, D = case(name)
    T = torch.DoubleTensor
    np.random.seed(5)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.RRF_Online(T(X), T(y), task, num_sampled_spectral=D)
        # same CPU RNG draws, in the same order, as the reference's constructor (A0 for this class)
        assert np.array_equal(m.gamma.cpu().numpy(), G[name + "_gamma0"])
        assert np.array_equal(m.w.cpu().numpy(), G[name + "_w0"])
        assert np.array_equal(m.eps.cpu().numpy(), G[name + "_eps"])
        pred, real, secs = m.online_learning()
    ref = G[name + "_pred"]
    assert len(pred) == len(ref) and len(real) == len(ref) and secs > 0
    if task == "cls":
        assert np.array_equal(pred, ref)
    else:
        np.testing.assert_allclose(pred, ref, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(m.w.cpu().numpy(), G[name + "_w"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m.gamma.cpu().numpy(), G[name + "_gamma"], rtol=1e-9, atol=1e-12)
