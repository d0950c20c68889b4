"""AFMAdam: the CUDA forward / backward against oracle/afm.py, the PyTorch definition of the AFM
paper's model (the reference's afm_adam.py cannot run, so parity is unpinned by the reference and the bar is fp32
agreement with autograd: logits 1e-5, gradients 1e-4 relative to their scale; exp/softmax orders are not mirrored)."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(sizes, k, A, lr, seed):
    import fm_for_online_recommendation_b200 as pkg
    from oracle.afm import AFMTorch
    torch.manual_seed(seed)
    m = pkg.AFMAdam(sizes, embedding_size=k, attention_size=A, n=lr)
    o = AFMTorch(sizes, k, A, n=lr)
    with torch.no_grad():
        m._table.mul_(0.3)
        t = m._table.cpu()
        o.V.copy_(t[:, :k]); o.w1.copy_(t[:, k]); o.bias.copy_(m.bias.cpu())
        att = m._att.cpu()
        o.W.copy_(att[:A * k].view(A, k)); o.c.copy_(att[A * k:A * k + A])
        o.H.copy_(att[A * k + A:A * k + 2 * A]); o.P.copy_(att[A * k + 2 * A:])
    return m, o


@pytest.mark.parametrize("sizes,k,A,B", [([5, 7, 3, 11], 6, 4, 64), ([63, 113, 9, 1457, 3, 305, 19, 632], 10, 4, 257),
                                          ([9] * 39, 10, 4, 128)])
def test_forward_and_gradients_match_autograd(sizes, k, A, B):
    m, o = _pair(sizes, k, A, 1e-3, 0)
    rng = np.random.RandomState(1)
    Xi = np.stack([rng.randint(0, fs, B) for fs in sizes], 1)
    Xv = rng.uniform(0.5, 1.5, Xi.shape).astype(np.float32)
    Y = (rng.uniform(size=B) < 0.4).astype(np.float32)
    z = m.forward(Xi, Xv).cpu().numpy()
    with torch.no_grad():
        zr = o.forward(Xi, Xv).numpy()
    np.testing.assert_allclose(z, zr, rtol=1e-5, atol=1e-5)
    assert np.array_equal(m.predict(Xi, Xv), zr > 0) or np.mean(m.predict(Xi, Xv) == (zr > 0)) > 0.99
    before = m._table.cpu().numpy().copy()
    PD = A * k + 2 * A + k
    gd = torch.zeros(PD, device="cuda")
    loss = m.update_embedding(Xi, Xv, Y, _grads_out=gd).item()
    lr_, g = o.step(Xi, Xv, Y)
    assert abs(loss - lr_) <= 1e-5 * max(1.0, abs(lr_))
    dense = np.concatenate([g["W"].reshape(-1), g["c"], g["H"], g["P"]])
    np.testing.assert_allclose(gd.cpu().numpy(), dense, rtol=1e-4, atol=1e-4 * np.abs(dense).max())
    # embedding rows: the sign step moves a coordinate by ~lr in the direction of -grad; compare on coordinates whose
    # gradient is clearly non-zero (a near-zero gradient may round to either sign)
    after = m._table.cpu().numpy()
    gV = g["V"]; gw = g["w1"]
    big = np.abs(gV) > 1e-4 * np.abs(gV).max()
    step = after[:, :k] - before[:, :k]
    assert np.all(np.sign(step[big]) == -np.sign(gV[big]))
    # fresh-Adam step: lr * |g| / (|g| + 1e-8); the gradient itself agrees with autograd to ~1e-4 of its scale
    clear = np.abs(gV) > 1e-2 * np.abs(gV).max()
    ag = np.abs(gV[clear]).astype(np.float64)
    assert np.allclose(np.abs(step[clear]), 1e-3 * ag / (ag + 1e-8), rtol=5e-2)
    untouched = np.ones(len(before), bool); untouched[np.unique(Xi + np.concatenate([[0], np.cumsum(sizes)])[:-1])] = False
    assert np.array_equal(after[untouched], before[untouched])
    bigw = np.abs(gw) > 1e-4 * np.abs(gw).max()
    assert np.all(np.sign((after[:, k] - before[:, k])[bigw]) == -np.sign(gw[bigw]))


def test_training_reduces_the_loss_and_the_model_pickles():
    sizes, k, A = [20, 30, 5, 40], 6, 4
    m, _ = _pair(sizes, k, A, 3e-3, 2)
    rng = np.random.RandomState(3)
    n = 2048
    Xi = np.stack([rng.randint(0, fs, n) for fs in sizes], 1)
    Xv = np.ones(Xi.shape, np.float32)
    teacher = rng.standard_normal(sum(sizes))
    off = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    Y = (teacher[Xi + off].sum(1) > 0).astype(np.float32)
    m.n_epochs, m.batch_size = 12, 256
    losses = m.fit(Xi, Xv, Y)
    assert losses[-1] < losses[0] - 0.05
    back = pickle.loads(pickle.dumps(m))
    assert str(back) == str(m) and np.array_equal(back.predict(Xi[:100], Xv[:100]), m.predict(Xi[:100], Xv[:100]))
