"""Pins the oracle's ATen mirrors (oracle/oracle_math.h group 1) on torch itself, bit for bit.

These run in the build container (torch 2.11 CPU with AVX-512, MKL, glibc 2.39 -- the host that generated
tests/golden).  torch.sigmoid's vector/scalar split depends on the CPU's vector width, and MKL's vsSqrt on its
CPU dispatch, so the torch comparisons skip on a host that is not AVX-512; the libm comparison always runs.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle.deep import lib

avx512 = pytest.mark.skipif(torch.backends.cpu.get_cpu_capability() != "AVX512",
                            reason="the mirrors restate torch's AVX-512 dispatch")


def run(fn, x):
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    f = getattr(L, fn)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    f(x.ctypes.data, y.ctypes.data, x.size)
    return y


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def logits(n, seed):
    rng = np.random.RandomState(seed)
    return (rng.standard_normal(n) * rng.choice([0.05, 1, 5, 20, 60, 120], n)).astype(np.float32)


@avx512
@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 250, 256, 4096, 8192, 8200, 32767])
def test_sigmoid_mirror_equals_torch_sigmoid(n):
    """vector body = Sleef expf u10, the last n % 32 elements = glibc expf (position-dependent bits)."""
    torch.set_num_threads(1)
    for seed in range(8):
        x = logits(n, seed)
        assert np.array_equal(bits(run("orc_vec_sigmoid", x)), bits(torch.sigmoid(torch.from_numpy(x)).numpy())), n
    # sigmoid of a sigmoid (FMAdam.fit's loss, fm_adam.py:80)
    p = torch.sigmoid(torch.from_numpy(logits(n, 99))).numpy()
    assert np.array_equal(bits(run("orc_vec_sigmoid", p)), bits(torch.sigmoid(torch.from_numpy(p)).numpy()))


@avx512
def test_log_sigmoid_mirror_equals_torch():
    torch.set_num_threads(1)
    for n in (1, 7, 33, 8192, 1 << 20):
        x = logits(n, n)
        want = torch.nn.functional.logsigmoid(torch.from_numpy(x)).numpy()
        assert np.array_equal(bits(run("orc_vec_log_sigmoid", x)), bits(want)), n


def test_sqrt_mirror_equals_torch_sqrt_on_every_mantissa():
    """Tensor.sqrt() is MKL's vsSqrt: 2^24 (mantissa, exponent parity) combinations at three exponents, all
    denormals, and the special values."""
    torch.set_num_threads(8)
    probe = np.array([float.fromhex("0x1.0824360000000p-22")], np.float32)
    if float(torch.from_numpy(probe).sqrt()[0]) == float(np.sqrt(probe)[0]):
        pytest.skip("this host's MKL returns the correctly rounded root (not the Intel VRSQRT14 path)")
    base = np.arange(1 << 24, dtype=np.uint32)
    for e0 in (1, 127, 252):
        x = ((np.uint32(e0) << np.uint32(23)) + base).view(np.float32)
        assert np.array_equal(bits(run("orc_vec_sqrt_mkl", x)), bits(torch.from_numpy(x).sqrt().numpy())), e0
    x = np.arange(0, 1 << 23, dtype=np.uint32).view(np.float32)
    assert np.array_equal(bits(run("orc_vec_sqrt_mkl", x)), bits(torch.from_numpy(x).sqrt().numpy()))
    x = np.array([0.0, -0.0, np.inf, 1.0, 4.0, 0.25, 3.4e38, 1e-45], np.float32)
    assert np.array_equal(bits(run("orc_vec_sqrt_mkl", x)), bits(torch.from_numpy(x).sqrt().numpy()))
    frac = float((bits(run("orc_vec_sqrt_mkl", base.view(np.float32)[1 << 23:])) !=
                  bits(np.sqrt(base.view(np.float32)[1 << 23:]))).mean())
    assert 0.003 < frac < 0.01      # ~0.6 % of inputs are one ulp below the correctly rounded root


def test_expf_glibc_mirror_equals_libm_expf():
    """sampled here (50 M inputs); oracle/verify_math.c is the exhaustive 2^32 run."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-110, 95, 200000), rng.standard_normal(100000) * 1e-3,
                        [0.0, -0.0, 88.72, 88.73, -103.9, -103.98, -103.3, -87.4, 1e-30, np.inf, -np.inf,
                         float.fromhex("0x1.04845ep+5"), float.fromhex("-0x1.f8cbb2p+5")]]).astype(np.float32)
    got = run("orc_vec_expf_glibc", x)
    want = np.array([libm.expf(float(v)) for v in x], np.float32)
    assert np.array_equal(bits(got), bits(want))


@avx512
def test_adam_first_step_equals_torch_optimizer_on_small_parameters():
    """the case round 1 missed: |p| ~ 0.2 exposes the one-ulp-low MKL square root in the denominator"""
    rng = np.random.RandomState(3)
    n = 1682 * 10
    p0 = (rng.standard_normal(n) * 0.2).astype(np.float32)
    g = (rng.standard_normal(n) * np.exp(rng.uniform(-45, 0, n))).astype(np.float32)
    g[::11] = 0.0
    for lr in (1e-4, 1e-3, 1e-2):
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        p.grad = torch.from_numpy(g.copy())
        torch.optim.Adam([p], lr=torch.nn.Parameter(torch.tensor(lr), requires_grad=False)).step()
        mine = p0.copy()
        L = lib()
        L.orc_update_dense(mine.ctypes.data_as(C.POINTER(C.c_float)), g.ctypes.data_as(C.POINTER(C.c_float)), n,
                           C.c_float(np.float32(lr)), 0)
        assert np.array_equal(bits(mine), bits(p.detach().numpy())), lr


def test_sgd_mode_equals_torch_optim_sgd():
    """update mode 1 pinned on torch: torch.optim.SGD's param.add_(grad, alpha=-lr) is one fused multiply-add (vector
    body and scalar tail alike); legacy notebooks, .ipynb_checkpoints/Online FM-checkpoint.ipynb cell 1."""
    rng = np.random.RandomState(0)
    for n in (7, 4096, 4099):
        p0 = rng.standard_normal(n).astype(np.float32)
        g = (rng.standard_normal(n) * np.exp(rng.uniform(-20, 2, n))).astype(np.float32)
        for lr in (0.01, 0.003):
            p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
            p.grad = torch.from_numpy(g.copy())
            torch.optim.SGD([p], lr=lr).step()
            mine = p0.copy()
            lib().orc_update_dense(mine.ctypes.data_as(C.POINTER(C.c_float)), g.ctypes.data_as(C.POINTER(C.c_float)), n,
                                   C.c_float(np.float32(lr)), 1)
            assert np.array_equal(bits(mine), bits(p.detach().numpy())), (n, lr)


def test_ftrl_proximal_restatement_follows_mcmahans_closed_form():
    """update mode 2: the fp32 restatement against a float64 evaluation of the per-coordinate
    FTRL-Proximal update over 50 steps of random gradients (the reference never exercises it: no torch oracle exists)."""
    rng = np.random.RandomState(1)
    n = 2000
    alpha, beta, l1, l2 = 0.05, 1.0, 0.02, 0.01
    w = (rng.standard_normal(n) * 0.1).astype(np.float32)
    z = (-(w.astype(np.float64) * (beta / alpha + l2)) - np.sign(w) * l1).astype(np.float32)
    nacc = np.zeros(n, np.float32)
    w64, z64, n64 = w.astype(np.float64), z.astype(np.float64), nacc.astype(np.float64)
    L = lib()
    fp = C.POINTER(C.c_float)
    L.orc_vec_ftrl.restype = None
    L.orc_vec_ftrl.argtypes = [fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float]
    for _ in range(50):
        g = (rng.standard_normal(n) * np.exp(rng.uniform(-6, 0, n))).astype(np.float32)
        L.orc_vec_ftrl(w.ctypes.data_as(fp), g.ctypes.data_as(fp), z.ctypes.data_as(fp), nacc.ctypes.data_as(fp), n,
                       alpha, beta, l1, l2)
        g64 = g.astype(np.float64)
        nn = n64 + g64 * g64
        sigma = (np.sqrt(nn) - np.sqrt(n64)) / alpha
        z64 = z64 + g64 - sigma * w64
        n64 = nn
        w64 = np.where(np.abs(z64) <= l1, 0.0, -(z64 - np.sign(z64) * l1) / ((beta + np.sqrt(n64)) / alpha + l2))
    # coordinates sitting on the L1 threshold may differ in the zero / non-zero decision; everything else agrees
    both = (w != 0) & (w64 != 0)
    assert both.mean() > 0.5 and (w == 0).any()
    np.testing.assert_allclose(w[both], w64[both], rtol=2e-3, atol=2e-5)
    assert ((w == 0) != (w64 == 0)).mean() < 0.01
