"""GPU parity of the tower models (A4-A8): DeepFM / NFM `fit`, hedge-backprop `fit`, predict and the
persistent per-example `run_experiment`, CUDA vs the CPU oracle (bit-exact: the SIMT tower accumulates
every contraction left to right with one FMA per term, like oracle/fm_oracle.c), plus replays of the
reference-generated golden fixtures (tolerance 1e-5, see tests/test_oracle_vs_golden.py)."""
import pickle

import numpy as np
import pytest
import torch

from _util import GOLDEN_DEEP, auc, load_golden, rel_err, rmse, synth
from test_gpu_fm import _pair, assert_same_params, pull, push

pytestmark = pytest.mark.gpu

SMALL = [7, 5, 11, 3, 13, 4]


def same_all(m, orc):
    assert_same_params(m, orc, exact=True)
    p = pull(m)
    if orc.L:
        assert np.array_equal(p["mlp"], orc.mlp), int((p["mlp"] != orc.mlp).sum())
    if hasattr(m, "alpha"):
        assert np.array_equal(p["alpha"], orc.alpha[:orc.L])


@pytest.mark.parametrize("kind,k,L,H,B", [("DeepFMAdam", 10, 3, 16, 64), ("NFMAdam", 64, 1, 64, 32),
                                          ("DeepFMAdam", 10, 5, 10, 250), ("NFMAdam", 8, 2, 24, 1),
                                          ("DeepFMAdam", 10, 3, 400, 96)])
def test_tower_forward_fit_update_bit_exact(kind, k, L, H, B):
    m, orc = _pair(kind, SMALL, k, L, H, lr=0.001, scale=0.2)
    for step in range(3):
        Xi, Xv, Y = synth(SMALL, B, 20 + step, real_xv=True, zipf=(step == 1))
        assert np.array_equal(m.forward(Xi, Xv).cpu().numpy(), orc.forward(Xi, Xv))
        assert np.array_equal(m.predict(Xi, Xv), orc.predict(Xi, Xv))
        m.fit(Xi, Xv, Y)
        orc.fit(Xi, Xv, Y)
        same_all(m, orc)
        got = float(m.update_embedding(Xi, Xv, Y).cpu())
        assert np.float32(got) == np.float32(orc.update_embedding(Xi, Xv, Y))
        same_all(m, orc)


@pytest.mark.parametrize("kind,bs", [("DeepFMOnn", 1), ("NFMOnn", 1), ("NFMOnn", 8), ("DeepFMOnn", 5)])
def test_hedge_fit_bit_exact(kind, bs):
    m, orc = _pair(kind, SMALL, 10, 5, 10, lr=0.01, scale=0.2, batch_size=bs)
    for step in range(4):
        Xi, Xv, Y = synth(SMALL, bs, 30 + step, real_xv=True)
        out, layers = m.forward(Xi, Xv)
        ro, rl = orc.forward(Xi, Xv)
        assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(layers.cpu().numpy(), rl)
        m.fit(Xi, Xv, Y)
        orc.fit(Xi, Xv, Y)
        same_all(m, orc)
    with pytest.raises(RuntimeError):
        m.fit(*synth(SMALL, bs + 1, 1))


@pytest.mark.parametrize("kind,L,H", [("FMAdam", 0, 0), ("DeepFMAdam", 3, 16), ("NFMAdam", 2, 10), ("DeepFMOnn", 5, 10),
                                      ("NFMOnn", 5, 10)])
def test_run_experiment_persistent_kernel_bit_exact(kind, L, H):
    m, orc = _pair(kind, SMALL, 10, L, H, lr=0.001, scale=0.2)
    Xi, Xv, Y = synth(SMALL, 300, 77, real_xv=True)
    _, acc, roc, conf = m.run_experiment(Xi.tolist(), Xv.tolist(), [int(v) for v in Y])
    oconf, opreds = orc.run_experiment(Xi, Xv, Y)
    assert conf == oconf
    assert np.array_equal(m._last_online_preds, opreds)
    same_all(m, orc)
    assert abs(acc - (conf["tp"] + conf["tn"]) / 300 * 100) < 1e-9 and set(roc) == {"tpr", "fpr"}


@pytest.mark.parametrize("name", list(GOLDEN_DEEP))
def test_golden_replay_against_reference_outputs(name):
    """the CUDA classes replay what the reference itself did (fixtures from tests/golden/make_golden.py)."""
    import fm_for_online_recommendation_b200 as pkg
    kind, kw = GOLDEN_DEEP[name]
    g = load_golden(name)
    ckw = dict(embedding_size=kw["k"], n=float(g["lr"]))
    if kind != "FMAdam":
        ckw.update(num_hidden_layers=kw["L"], neuron_per_hidden_layer=kw["H"])
    if "batch_size" in kw:
        ckw["batch_size"] = kw["batch_size"]
    m = getattr(pkg, kind)(g["feature_sizes"].tolist(), **ckw)

    class O:  # parameter carrier with the oracle's attribute names
        pass
    o = O()
    o.k, o.L = kw["k"], kw.get("L", 0)
    o.V, o.w1, o.bias, o.mlp = g["init_V"], g["init_w1"], g["init_bias"], g["init_mlp"]
    o.alpha = g.get("init_alpha", np.zeros(1, np.float32))
    push(m, o)
    f = m.forward(g["Xi"].tolist(), g["Xv"].tolist())
    f0 = (f[0] if isinstance(f, tuple) else f).cpu().numpy()
    np.testing.assert_allclose(f0, g["fwd0"], rtol=1e-5, atol=1e-5)
    if "fwd_fm0" in g:
        assert np.array_equal(m.forward_fm(g["Xi"], g["Xv"]).cpu().numpy(), g["fwd_fm0"])  # logits: bit-exact
    assert np.array_equal(m.predict(g["Xi"].tolist(), g["Xv"].tolist()), g["pred0"].reshape(-1))
    losses = [float(m.update_embedding(g["ue_Xi"][s].tolist(), g["ue_Xv"][s].tolist(), g["ue_Y"][s].tolist()).cpu())
              for s in range(int(g["steps"]))]
    np.testing.assert_allclose(losses, g["ue_loss"], rtol=1e-5)
    for s in range(int(g["steps"])):
        m.fit(g["fit_Xi"][s].tolist(), g["fit_Xv"][s].tolist(), g["fit_Y"][s].tolist())
    p = pull(m)
    assert rel_err(p["V"], g["after_fit_V"]) <= 1e-5 and rel_err(p["w1"], g["after_fit_w1"]) <= 1e-5
    assert rel_err(p["bias"], g["after_fit_bias"]) <= 1e-5
    # north_star: AUC and RMSE identical to 4 decimal places (scores after the golden training trajectory)
    f = m.forward(g["Xi"].tolist(), g["Xv"].tolist())
    z = (f[0] if isinstance(f, tuple) else f).cpu().numpy().astype(np.float64)
    zr, y = np.asarray(g["fwd1"], np.float64), np.asarray(g["Y"]).reshape(-1)
    if len(np.unique(y > 0)) == 2:
        assert round(auc(z, y), 4) == round(auc(zr, y), 4)
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))
    pz, pr = (z, zr) if isinstance(f, tuple) else (sig(z), sig(zr))
    assert round(rmse(pz, y), 4) == round(rmse(pr, y), 4)
    if o.L:
        assert rel_err(p["mlp"], g["after_fit_mlp"]) <= 1e-5
    if "on_Xi" in g:
        _, _, _, conf = m.run_experiment(g["on_Xi"].tolist(), g["on_Xv"].tolist(), [int(v) for v in g["on_Y"]])
        assert [conf["tp"], conf["fp"], conf["tn"], conf["fn"]] == g["on_conf"].tolist()
        p = pull(m)
        assert rel_err(p["V"], g["after_on_V"]) <= 1e-5


def test_pickle_roundtrip_and_str():
    import fm_for_online_recommendation_b200 as pkg
    torch.manual_seed(3)
    m = pkg.DeepFMOnn(SMALL, embedding_size=10, num_hidden_layers=5, neuron_per_hidden_layer=10, n=1e-4)
    assert str(m).split("-")[0] == "DeepFMOnn"  # main_experiment.py:86
    Xi, Xv, Y = synth(SMALL, 16, 5)
    m.update_embedding(Xi, Xv, Y)
    m2 = pickle.loads(pickle.dumps(m))          # main_experiment.py:160-162
    assert torch.equal(m2._table, m._table) and torch.equal(m2._mlp, m._mlp) and torch.equal(m2.alpha, m.alpha)
    assert np.array_equal(m2.predict(Xi, Xv), m.predict(Xi, Xv))
    names = [n for n, _ in m.named_parameters()]
    assert "first_order_embeddings.0.weight" in names and "hidden_layers.4.bias" in names
    assert m.second_order_embeddings[2].weight.shape == (SMALL[2], 10)


def test_same_seed_same_init_as_reference_rng_order():
    """parameters are drawn on the CPU in the reference's construction order (fm_adam.py:26-32)."""
    import fm_for_online_recommendation_b200 as pkg
    g = load_golden("fm_cfg1")
    torch.manual_seed(0)
    m = pkg.FMAdam([943, 1682], embedding_size=10, n=0.01)
    p = pull(m)
    assert np.array_equal(p["V"], g["init_V"]) and np.array_equal(p["w1"], g["init_w1"])


def test_tensor_core_gemm_3xtf32_within_tolerance():
    """csrc/gemm_tc.cu: tcgen05.mma kind::tf32 x3 with fp32 TMEM accumulation vs an fp64 reference.  Tolerance 1e-5
    of max|C| (BASELINE.json's bound); K > 1024 is split across CTAs so one TMEM pass never accumulates more."""
    import ctypes as C
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    torch.manual_seed(0)
    P = lambda t: C.c_void_p(t.data_ptr())
    for (M, N, K) in [(128, 128, 32), (300, 400, 400), (8192, 400, 10), (1000, 72, 8192), (77, 209, 133), (1, 1, 1),
                      (129, 417, 33)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        Cc = torch.full((M, N), float("nan"), device="cuda")
        assert lib.fmb_gemm_tc_nt(P(A), P(B), P(Cc), M, N, K, None) == 0
        torch.cuda.synchronize()
        assert lib.fmb_gemm_tc_error() == 0
        ref = (A.double() @ B.double().t())
        rel = ((Cc.double() - ref).abs().max() / ref.abs().max()).item()
        assert rel < 1e-5, (M, N, K, rel)


def test_tensor_core_gemm_strided_forms_and_epilogues():
    """The three tower products as mlp.cu issues them: forward NT + bias + relu, dX NN + relu mask, dW TN with
    K = batch (split-K) + column sums (the bias gradient); rows/columns outside the tile edges untouched."""
    import ctypes as C
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    torch.manual_seed(1)
    P = lambda t: C.c_void_p(t.data_ptr())
    for (Bt, H, nin) in [(8192, 400, 400), (3000, 72, 10), (515, 130, 77)]:
        gp = torch.randn(Bt, H, device="cuda")
        x = torch.randn(Bt, nin, device="cuda")
        W = torch.randn(H, nin, device="cuda")
        bias = torch.randn(H, device="cuda")
        mask = torch.randn(Bt, nin, device="cuda")
        y = torch.zeros(Bt, H, device="cuda")
        assert lib.fmb_gemm_tc_strided(P(x), nin, 1, P(W), 1, nin, P(y), H, Bt, H, nin, 1, P(bias), None, 0, None,
                                       None) == 0
        ref = torch.relu(x.double() @ W.double().t() + bias.double())
        assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
        gx = torch.zeros(Bt, nin, device="cuda")
        assert lib.fmb_gemm_tc_strided(P(gp), H, 1, P(W), nin, 1, P(gx), nin, Bt, nin, H, 2, None, P(mask), nin, None,
                                       None) == 0
        ref = (gp.double() @ W.double()) * (mask > 0)
        assert ((gx.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
        gW = torch.zeros(H, nin, device="cuda")
        gc = torch.zeros(H, device="cuda")
        assert lib.fmb_gemm_tc_strided(P(gp), 1, H, P(x), nin, 1, P(gW), nin, H, nin, Bt, 0, None, None, 0, P(gc),
                                       None) == 0
        ref = gp.double().t() @ x.double()
        assert ((gW.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
        refc = gp.double().sum(0)
        assert ((gc.double() - refc).abs().max() / refc.abs().max()).item() < 1e-5
        gW2 = torch.zeros(H, nin, device="cuda")   # split-K partials are added in a fixed order: run-to-run identical
        assert lib.fmb_gemm_tc_strided(P(gp), 1, H, P(x), nin, 1, P(gW2), nin, H, nin, Bt, 0, None, None, 0, P(gc),
                                       None) == 0
        torch.cuda.synchronize()
        assert torch.equal(gW, gW2)
    assert lib.fmb_gemm_tc_error() == 0


def test_tower_backward_tensor_cores_vs_exact_simt_same_activations():
    """fmb_mlp_backward at the cfg4 shape on the SAME saved activations: tensor-core gradients within 1e-5 of the
    exact SIMT ones (relative to the largest gradient)."""
    import ctypes as C
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    torch.manual_seed(2)
    P = lambda t: C.c_void_p(t.data_ptr())
    B, k, L, H = 8192, 10, 3, 400
    bi = torch.randn(B, k, device="cuda")
    n = lib.fmb_mlp_numel(k, L, H)
    mlp = (torch.rand(n, device="cuda") - 0.5) * 0.1
    act = torch.empty(L, B, H, device="cuda")
    gtop = torch.randn(B, device="cuda")
    wsb = lib.fmb_mlp_bwd_workspace_bytes(B, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lib.fmb_set_tensor_cores(0)
    assert lib.fmb_mlp_forward(P(bi), k, P(mlp), B, k, L, H, P(act), None, None) == 0
    out = []
    for tc in (0, 1):
        lib.fmb_set_tensor_cores(tc)
        gmlp = torch.zeros(n, device="cuda")
        gbi = torch.zeros(B, k, device="cuda")
        assert lib.fmb_mlp_backward(P(bi), k, P(mlp), P(act), P(gtop), L - 1, B, k, L, H, P(gmlp), P(gbi), k, P(ws),
                                    wsb, None) == 0
        torch.cuda.synchronize()
        out.append((gmlp, gbi))
    lib.fmb_set_tensor_cores(1)
    assert lib.fmb_gemm_tc_error() == 0
    for a, b in zip(out[0], out[1]):
        assert ((a - b).abs().max() / a.abs().max()).item() < 1e-5


def test_cfg4_tower_fit_tensor_cores_vs_exact_simt():
    """BASELINE configs[3] shape (B = 8192, k = 10, 400-400-400): one DeepFMAdam.fit with the tcgen05 tower must
    agree with the exact SIMT tower (which is bit-identical to the oracle) within 1e-5 on the logits, and the
    sign-step updates of the tables must agree on >= 99 % of the coordinates and differ by at most 2 lr."""
    import fm_for_online_recommendation_b200 as pkg
    from test_gpu_fm import CRITEO
    lib = pkg.require_cuda()
    B = 8192
    Xi, Xv, Y = synth(CRITEO, B, 7)
    outs = []
    for tc in (1, 0):
        lib.fmb_set_tensor_cores(tc)
        torch.manual_seed(5)
        m = pkg.DeepFMAdam(CRITEO, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
        with torch.no_grad():
            m._table[:, :11].mul_(0.1)
        e = m.encode(Xi, Xv, Y)
        z0 = m.forward(e, None).clone()
        m.fit(e, None, None)
        outs.append((z0.cpu().numpy(), m._table.cpu().numpy().copy(), m._mlp.cpu().numpy().copy()))
    lib.fmb_set_tensor_cores(1)
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-5)
    # the first Adam step is a sign step (+-lr whatever |g| is), so gradient noise at the 1e-6 level flips the
    # coordinates whose gradient cancels to ~0; everything else is identical, and a flip moves a value by 2 lr
    same = (outs[0][1] == outs[1][1]).mean()
    assert same > 0.99, same
    assert np.abs(outs[0][1] - outs[1][1]).max() <= 2.1e-4
    assert np.abs(outs[0][2] - outs[1][2]).max() <= 2.1e-4   # an MLP weight moves by +-lr at most


@pytest.mark.parametrize("kind", ["DeepFMOnn", "NFMOnn"])
def test_cfg4_hedge_fit_single_pass_vs_L_passes(kind, monkeypatch):
    """BASELINE configs[3] shape (B = 8192, 400-400-400 tower): the single-pass hedge backward (alpha_i dL_i/dhead_i injected at
    every head, fmb_mlp_backward_hedge) against the reference's structure (one backward pass per head, deepfm_onn.py:127-141)
    -- the same gradient sums up to fp32 rounding.  After ONE fit: tower weights within 2e-6 relative, alpha within 1e-6.
    After three fits (the second and third graph-replayed) the two trajectories have drifted by rounding amplified through
    ReLU boundaries (measured 5e-7 absolute = 4e-5 of the accumulated update, tools/diag_hedge.py): bounded at 1e-3 of the
    accumulated update.  The tables are untouched (hedge fit does not train them)."""
    import fm_for_online_recommendation_b200 as pkg
    from test_gpu_fm import CRITEO
    B = 8192
    outs = []
    for single in ("1", "0"):
        monkeypatch.setenv("FMB_HEDGE_SINGLE", single)
        torch.manual_seed(5)
        kw = dict(embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-2, batch_size=B)
        m = getattr(pkg, kind)(CRITEO, **kw)
        with torch.no_grad():
            m._table[:, :11].mul_(0.05)
        t0 = m._table.clone()
        w0 = m._mlp.cpu().numpy().copy()
        snaps = []
        for step in range(3):
            Xi, Xv, Y = synth(CRITEO, B, 70 + step)
            m.fit(m.encode(Xi, Xv, Y), None, None)
            snaps.append((m._mlp.cpu().numpy().copy(), m.alpha.detach().cpu().numpy().copy()))
        assert torch.equal(t0, m._table)
        outs.append(snaps)
    one, three = (outs[0][0], outs[1][0]), (outs[0][2], outs[1][2])
    assert np.abs(one[1][0] - w0).max() > 1e-4                      # the step did move the tower
    assert rel_err(one[0][0], one[1][0]) <= 2e-6, rel_err(one[0][0], one[1][0])
    np.testing.assert_allclose(one[0][1], one[1][1], rtol=1e-6, atol=1e-7)
    moved = np.abs(three[1][0] - w0).max()
    assert np.abs(three[0][0] - three[1][0]).max() <= 1e-3 * moved
    np.testing.assert_allclose(three[0][1], three[1][1], rtol=1e-5, atol=1e-6)
