"""Pins the CPU oracle (oracle/fm_oracle.c) against outputs of the reference itself.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports
haan6/fm-for-online-recommendation from /root/reference and runs its own classes
(models/models_online_deep/*.py, use_cuda=False) on seeded synthetic inputs.

Bars (BASELINE.json north_star): first-order / Bi-Interaction values and logits bit-exact;
losses and updated weights within 1e-5 relative; online confusion counts identical.
The only arithmetic the oracle cannot mirror bit for bit is torch's sigmoid/log (ISA- and
batch-size-dependent) and MKL's GEMM order, hence <= 1e-5 rather than 0 on those.
"""
import numpy as np
import pytest

from _util import GOLDEN_DEEP, auc, load_golden, rel_err, rmse
from oracle.deep import OracleDeep, lib

TOL = 1e-5


def make(name):
    kind, kw = GOLDEN_DEEP[name]
    g = load_golden(name)
    m = OracleDeep(kind, g["feature_sizes"].tolist(), lr=float(g["lr"]), **kw)
    return m, g


def load(m, g, prefix):
    m.w1[:] = g[prefix + "w1"]
    m.V[:] = g[prefix + "V"]
    m.bias[:] = g[prefix + "bias"]
    if m.L:
        m.mlp[:] = g[prefix + "mlp"]
    if prefix + "alpha" in g:
        m.alpha[:] = g[prefix + "alpha"]


def check_params(m, g, prefix, tol=TOL):
    assert rel_err(m.w1, g[prefix + "w1"]) <= tol
    assert rel_err(m.V, g[prefix + "V"]) <= tol
    assert rel_err(m.bias, g[prefix + "bias"]) <= tol
    if m.L:
        assert rel_err(m.mlp, g[prefix + "mlp"]) <= tol
    if prefix + "alpha" in g:
        assert rel_err(m.alpha, g[prefix + "alpha"]) <= tol


@pytest.mark.parametrize("name", list(GOLDEN_DEEP))
def test_forward_bit_exact(name):
    m, g = make(name)
    load(m, g, "init_")
    p = m.fm_parts(g["Xi"], g["Xv"])
    if "first0" in g:
        assert np.array_equal(p["first"], g["first0"])       # A1 first_order
        assert np.array_equal(p["bi"], g["second0"])         # A2 second_order / Bi-Interaction
        assert np.array_equal(p["z_fm"], g["fwd_fm0"])       # A3 forward_fm (ATen row-sum order mirrored)
    f = m.forward(g["Xi"], g["Xv"])
    f0 = f[0] if isinstance(f, tuple) else f
    if m.L == 0:
        assert np.array_equal(f0, g["fwd0"])
    else:  # MLP: MKL's sgemm order is not mirrored
        np.testing.assert_allclose(f0, g["fwd0"], rtol=TOL, atol=TOL)
    if isinstance(f, tuple):
        np.testing.assert_allclose(f[1], g["fwd0_layers"], rtol=TOL, atol=TOL)
    assert np.array_equal(m.predict(g["Xi"], g["Xv"]), g["pred0"].reshape(-1))


@pytest.mark.parametrize("name", list(GOLDEN_DEEP))
def test_update_embedding_and_fit_trajectory(name):
    m, g = make(name)
    load(m, g, "init_")
    losses = [m.update_embedding(g["ue_Xi"][s], g["ue_Xv"][s], g["ue_Y"][s]) for s in range(int(g["steps"]))]
    np.testing.assert_allclose(losses, g["ue_loss"], rtol=TOL)
    check_params(m, g, "after_ue_")
    for s in range(int(g["steps"])):
        m.fit(g["fit_Xi"][s], g["fit_Xv"][s], g["fit_Y"][s])
    check_params(m, g, "after_fit_")
    f = m.forward(g["Xi"], g["Xv"])
    f1 = f[0] if isinstance(f, tuple) else f
    np.testing.assert_allclose(f1, g["fwd1"], rtol=TOL, atol=TOL)


@pytest.mark.parametrize("name", [n for n in GOLDEN_DEEP if "on_Xi" in load_golden(n)])
def test_run_experiment(name):
    m, g = make(name)
    load(m, g, "after_fit_")
    conf, _ = m.run_experiment(g["on_Xi"], g["on_Xv"], g["on_Y"])
    assert [conf["tp"], conf["fp"], conf["tn"], conf["fn"]] == g["on_conf"].tolist()
    check_params(m, g, "after_on_")


def test_teacher_forced_single_step_mostly_bit_exact():
    """One step from identical weights: every row whose update is not within reach of the sigmoid's
    last-ulp difference must match the reference bit for bit (>= 99.5 % of the table)."""
    for name in ("fm_cfg1", "deepfm_raw"):
        m, g = make(name)
        load(m, g, "init_")
        m.update_embedding(g["ue_Xi"][0], g["ue_Xv"][0], g["ue_Y"][0])
        # golden holds the state after all `steps` updates; replay the rest too and compare
        for s in range(1, int(g["steps"])):
            m.update_embedding(g["ue_Xi"][s], g["ue_Xv"][s], g["ue_Y"][s])
        frac = float((m.V == g["after_ue_V"]).mean())
        assert frac >= 0.995, frac


def test_adam1_matches_torch_optimizer_bitwise():
    """A12: the fresh-state Adam step restated in C equals torch.optim.Adam bit for bit."""
    import ctypes as C
    import torch
    rng = np.random.RandomState(0)
    p0 = rng.standard_normal(4096).astype(np.float32)
    g = (rng.standard_normal(4096) * np.exp(rng.uniform(-40, 2, 4096))).astype(np.float32)
    g[::7] = 0.0
    for lr in (1e-4, 0.01, 0.003):
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        p.grad = torch.from_numpy(g.copy())
        n = torch.nn.Parameter(torch.tensor(lr), requires_grad=False)
        torch.optim.Adam([p], lr=n).step()
        mine = p0.copy()
        lib().orc_update_dense(mine.ctypes.data_as(C.POINTER(C.c_float)), g.ctypes.data_as(C.POINTER(C.c_float)),
                               mine.size, C.c_float(np.float32(lr)), 0)
        assert np.array_equal(mine, p.detach().numpy())


def test_sum_aten_matches_torch_sum_bitwise():
    import ctypes as C
    import torch
    torch.set_num_threads(1)
    rng = np.random.RandomState(1)
    for n in (1, 3, 5, 7, 8, 10, 39, 64, 400, 511, 512, 2500, 8192, 20000):
        x = (rng.standard_normal(n) * 10).astype(np.float32)
        want = torch.sum(torch.from_numpy(x)).item()
        got = lib().orc_sum_aten(x.ctypes.data_as(C.POINTER(C.c_float)), n)
        assert np.float32(got) == np.float32(want), n


def test_oracle_host_threads_do_not_change_results():
    """oracle/fm_oracle.c splits the batch step over ORC_THREADS host threads (samples in the forward pass, fields in
    the backward pass) for the bench's CPU baseline: every bit of the result must be the same as with one thread."""
    import os
    from oracle.deep import OracleDeep
    sizes = [7, 300, 4, 1000, 50, 9]
    outs = []
    for threads in ("1", "4"):
        os.environ["ORC_THREADS"] = threads
        orc = OracleDeep("DeepFMAdam", sizes, 6, 2, 8, lr=0.01, seed=3)
        rng = np.random.RandomState(5)
        off = np.concatenate([[0], np.cumsum(sizes)])[:-1]
        losses = []
        for _ in range(3):
            Xi = np.stack([rng.randint(0, fs, size=600) for fs in sizes], 1)
            Xv = rng.uniform(0.5, 1.5, size=Xi.shape).astype(np.float32)
            Y = (rng.uniform(size=600) < 0.4).astype(np.float32)
            losses.append(orc.update_embedding(Xi, Xv, Y))
        outs.append((np.array(losses, np.float32), orc.V.copy(), orc.w1.copy(), orc.bias.copy()))
        del off
    os.environ["ORC_THREADS"] = "1"
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", list(GOLDEN_DEEP))
def test_auc_and_rmse_identical_to_4_decimals(name):
    """BASELINE.json north_star: "AUC and RMSE must be identical to 4 decimal places".  After the golden training
    trajectory (update_embedding steps, then fit steps) the oracle's scores on the evaluation batch give the same
    AUC and the same RMSE (of sigmoid(z) against the 0/1 labels) as the reference's recorded scores."""
    m, g = make(name)
    load(m, g, "init_")
    for s in range(int(g["steps"])):
        m.update_embedding(g["ue_Xi"][s], g["ue_Xv"][s], g["ue_Y"][s])
    for s in range(int(g["steps"])):
        m.fit(g["fit_Xi"][s], g["fit_Xv"][s], g["fit_Y"][s])
    f = m.forward(g["Xi"], g["Xv"])
    z = np.asarray(f[0] if isinstance(f, tuple) else f, np.float64)
    zr = np.asarray(g["fwd1"], np.float64)
    y = np.asarray(g["Y"]).reshape(-1)
    if len(np.unique(y > 0)) == 2:
        assert round(auc(z, y), 4) == round(auc(zr, y), 4)
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))
    # ONN forward already returns probabilities; the Adam family returns logits
    pz, pr = (z, zr) if isinstance(f, tuple) else (sig(z), sig(zr))
    assert round(rmse(pz, y), 4) == round(rmse(pr, y), 4)
