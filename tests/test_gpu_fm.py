"""GPU parity tests of the FM hot path (A1-A3, sort/segments, A6) through the C ABI.
CUDA results are compared with the CPU oracle bit for bit (fp32 work mirrors ATen's op order)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from _util import synth

pytestmark = pytest.mark.gpu

CRITEO = [63, 113, 126, 51, 224, 148, 100, 79, 104, 9, 32, 57, 82, 1457, 555, 176373, 129683, 305, 19, 11887,
          632, 3, 41738, 5170, 175446, 3170, 27, 11356, 165602, 10, 4641, 2030, 4, 172761, 18, 15, 57903, 86, 44549]
FRAPPE = [957, 4082, 7, 7, 2, 3, 2, 9, 80, 233]


def _pair(kind, sizes, k, L=0, H=0, lr=0.01, seed=0, scale=1.0, **kw):
    """(product model on cuda, oracle) holding identical parameters."""
    import fm_for_online_recommendation_b200 as pkg
    from oracle.deep import OracleDeep
    orc = OracleDeep(kind, sizes, k, L, H, lr=lr, seed=seed, **kw)
    orc.w1 *= np.float32(scale)
    orc.V *= np.float32(scale)
    cls = getattr(pkg, kind)
    ckw = dict(embedding_size=k, n=lr)
    if kind != "FMAdam":
        ckw.update(num_hidden_layers=L, neuron_per_hidden_layer=H)
    if "batch_size" in kw:
        ckw["batch_size"] = kw["batch_size"]
    m = cls(sizes, **ckw)
    push(m, orc)
    return m, orc


def push(m, orc):
    k = orc.k
    with torch.no_grad():
        t = torch.zeros_like(m._table)
        t[:, :k] = torch.from_numpy(orc.V)
        t[:, k] = torch.from_numpy(orc.w1)
        m._table.copy_(t)
        m.bias.copy_(torch.from_numpy(orc.bias).reshape(m.bias.shape))
        if orc.L:
            m._mlp.copy_(torch.from_numpy(orc.mlp))
        if hasattr(m, "alpha"):
            m.alpha.copy_(torch.from_numpy(orc.alpha[:orc.L]))


def pull(m):
    k = m.embedding_size
    t = m._table.detach().cpu().numpy()
    out = dict(V=t[:, :k].copy(), w1=t[:, k].copy(), bias=m.bias.detach().cpu().numpy().reshape(1))
    if m._mlp is not None:
        out["mlp"] = m._mlp.detach().cpu().numpy()
    if hasattr(m, "alpha"):
        out["alpha"] = m.alpha.detach().cpu().numpy()
    return out


def assert_same_params(m, orc, exact=True, tol=1e-5):
    p = pull(m)
    for key in ("V", "w1", "bias"):
        a, b = p[key], getattr(orc, key)
        if exact:
            assert np.array_equal(a, b), (key, int((a != b).sum()), a.size)
        else:
            np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


@pytest.mark.parametrize("sizes,k,B,real_xv", [([943, 1682], 10, 256, False), (FRAPPE, 64, 300, True),
                                               (CRITEO, 10, 1000, False), ([5, 3], 4, 1, True),
                                               ([7, 5, 11, 3, 13, 4], 3, 33, True)])
def test_forward_bit_exact(sizes, k, B, real_xv):
    m, orc = _pair("FMAdam", sizes, k, scale=0.3)
    Xi, Xv, _ = synth(sizes, B, 5, real_xv=real_xv)
    ref = orc.fm_parts(Xi, Xv)
    assert np.array_equal(m.first_order(Xi, Xv).cpu().numpy(), ref["first"])
    assert np.array_equal(m.second_order(Xi, Xv).cpu().numpy(), ref["bi"])
    assert np.array_equal(m.forward_fm(Xi, Xv).cpu().numpy(), ref["z_fm"])
    assert np.array_equal(m.forward(Xi, Xv).cpu().numpy(), ref["z_fm"])
    assert np.array_equal(m.predict(Xi, Xv), orc.predict(Xi, Xv))


@pytest.mark.parametrize("N,bits,dist", [(1, 5, "u"), (31, 3, "u"), (2048, 8, "u"), (2049, 20, "u"),
                                         (100000, 25, "u"), (319488, 20, "zipf"), (70000, 1, "u"),
                                         (50000, 17, "const")])
def test_sort_and_segments_bit_exact(N, bits, dist):
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    rng = np.random.RandomState(N)
    if dist == "u":
        keys = rng.randint(0, 2 ** bits, size=N).astype(np.int32)
    elif dist == "zipf":
        keys = np.minimum(rng.zipf(1.2, size=N), 2 ** bits - 1).astype(np.int32)
    else:
        keys = np.full(N, 2 ** bits - 3, np.int32)
    d = torch.from_numpy(keys).cuda()
    wsb = lib.fmb_sort_workspace_bytes(N)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    sk = torch.empty(N, dtype=torch.int32, device="cuda")
    pm = torch.empty(N, dtype=torch.int32, device="cuda")
    seg = torch.empty(N + 1, dtype=torch.int32, device="cuda")
    nseg = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.fmb_sort_segment(p(d), N, bits, p(ws), wsb, p(sk), p(pm), p(seg), p(nseg), None)
    assert rc == 0, lib.fmb_last_error()
    torch.cuda.synchronize()
    order = np.argsort(keys, kind="stable").astype(np.int32)
    assert np.array_equal(pm.cpu().numpy(), order)            # stable permutation, bit-exact
    assert np.array_equal(sk.cpu().numpy(), keys[order])
    starts = np.flatnonzero(np.concatenate([[True], keys[order][1:] != keys[order][:-1]])).astype(np.int32)
    n = int(nseg.item())
    assert n == len(starts)
    assert np.array_equal(seg.cpu().numpy()[:n], starts) and int(seg[n].item()) == N


@pytest.mark.parametrize("B", [1, 33, 1000, 2048, 2049, 8192, 16384, 40001, 65536])
def test_sort_fields_matches_global_stable_sort(B):
    """per-field shared-memory sort == stable argsort of the flattened [B,F] id matrix (bit-exact)."""
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    sizes = [1, 3, 200, 257, 70000, 1269185, 2, 65536]
    F = len(sizes)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    rng = np.random.RandomState(B)
    ids = np.stack([rng.randint(0, fs, size=B) for fs in sizes], 1).astype(np.int32) + off[:-1][None, :]
    ids[:, 4] = off[4] + np.minimum(rng.zipf(1.2, size=B), sizes[4] - 1)
    d = torch.from_numpy(ids).cuda()
    doff = torch.from_numpy(off).cuda()
    sk = torch.empty(B * F, dtype=torch.int32, device="cuda")
    pm = torch.empty(B * F, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    assert lib.fmb_sort_fields(p(d), B, F, p(doff), p(sk), p(pm), None) == 0, lib.fmb_last_error()
    torch.cuda.synchronize()
    flat = ids.reshape(-1)
    order = np.argsort(flat, kind="stable").astype(np.int32)
    assert np.array_equal(pm.cpu().numpy(), order)
    assert np.array_equal(sk.cpu().numpy(), flat[order])
    # the sort's tail (posflag + run list), the stand-alone kernel behind the generic sort, and the sparse-field hash
    # pass (fields with >= 16*B rows are not sorted: only their multi-hit entries are placed) give what numpy gives
    from fm_for_online_recommendation_b200._lib import RunList
    exp_pf, exp_runs = _posflag_and_runs(flat, order)
    N = B * F
    nseg, cap = C.c_int(), C.c_int()
    lib.fmb_runlist_shape(B, F, C.byref(nseg), C.byref(cap))
    nseg, cap = nseg.value, cap.value
    skn, multi = flat[order], (exp_pf >> 31).astype(bool)
    for producer in ("sort_tail", "pos_flags", "sparse"):
        pf = torch.zeros(N, dtype=torch.int32, device="cuda")
        one = producer == "pos_flags"
        rl = torch.full((max(nseg * cap, N // 2 + 1), 4), -1, dtype=torch.int32, device="cuda")
        rc_ = torch.zeros(2 * nseg, dtype=torch.int32, device="cuda")
        desc = RunList(rl.data_ptr(), rc_.data_ptr(), 1 if one else nseg, N // 2 + 1 if one else cap)
        sk.zero_(); pm.zero_()
        if producer == "pos_flags":
            assert lib.fmb_sort_fields(p(d), B, F, p(doff), p(sk), p(pm), None) == 0
            assert lib.fmb_pos_flags_ex(p(sk), p(pm), N, p(pf), C.byref(desc), None) == 0, lib.fmb_last_error()
        else:
            assert lib.fmb_sort_fields_ex(p(d), B, F, p(doff), p(sk), p(pm), p(pf), C.byref(desc),
                                          1 if producer == "sparse" else 0, None) == 0, lib.fmb_last_error()
        torch.cuda.synchronize()
        counts = rc_.cpu().numpy()
        rln, cp_ = rl.cpu().numpy(), desc.seg_cap
        short = [rln[g * cp_: g * cp_ + counts[g]] for g in range(desc.nseg)]                                   # upwards from the start
        long_ = [rln[(g + 1) * cp_ - counts[desc.nseg + g]: (g + 1) * cp_] for g in range(desc.nseg)]            # downwards from the end
        got = np.concatenate(short + long_ + [np.zeros((0, 4), np.int32)])
        assert all(np.all(a[:, 3] == 0) for a in short) and all(np.all(a[:, 3] == 1) for a in long_)
        gpf = pf.cpu().numpy().view(np.uint32)
        gsk, gpm = sk.cpu().numpy(), pm.cpu().numpy()
        if producer != "sparse":
            assert np.array_equal(gpm, order) and np.array_equal(gsk, skn)
            assert np.array_equal(gpf, exp_pf), producer
            assert len(got) == len(exp_runs), producer
            assert np.array_equal(got[np.argsort(got[:, 0])], exp_runs), producer
            continue
        # sparse contract: same multi-hit flags; per field the multi-hit entries in (key, sample) order at the head of the
        # field's range, -1 behind them; every listed run is a run of the placed keys; dense fields as before
        assert np.array_equal(gpf >> 31, exp_pf >> 31)
        min_rows = lib.fmb_sort_fields_sparse_min_rows(B) or (1 << 40)
        for f in range(F):
            lo, hi = f * B, (f + 1) * B
            if sizes[f] < min_rows:
                assert np.array_equal(gsk[lo:hi], skn[lo:hi]) and np.array_equal(gpm[lo:hi], order[lo:hi])
                ent = order[lo:hi]
                assert np.array_equal(gpf[ent], exp_pf[ent])
                continue
            ms = multi[order[lo:hi]]                  # sorted positions of the field whose entry is multi-hit
            nm = int(ms.sum())
            assert np.array_equal(gsk[lo:lo + nm], skn[lo:hi][ms]) and np.all(gsk[lo + nm:hi] == -1)
            assert np.array_equal(gpm[lo:lo + nm], order[lo:hi][ms])
            assert np.array_equal(gpf[gpm[lo:lo + nm]] & 0x7fffffff, np.arange(lo, lo + nm, dtype=np.uint32))
        # run list == the runs of the placed keys
        placed = gsk.copy()
        same_prev = np.concatenate([[False], (placed[1:] == placed[:-1]) & (placed[1:] >= 0)])
        same_next = np.concatenate([(placed[:-1] == placed[1:]) & (placed[:-1] >= 0), [False]])
        starts = np.flatnonzero(same_next & ~same_prev)
        ends = np.flatnonzero(same_prev & ~same_next)
        exp2 = np.stack([starts, placed[starts], np.minimum(ends - starts + 1, 32), ends - starts + 1 >= 128], 1).astype(np.int32)
        assert len(got) == len(exp2)
        assert np.array_equal(got[np.argsort(got[:, 0])], exp2)


def _posflag_and_runs(flat, order):
    """numpy statement of radix_sort.cu sort_tail / fm_step.cu pos_flags_kernel"""
    sk = flat[order]
    N = len(sk)
    same_prev = np.concatenate([[False], sk[1:] == sk[:-1]])
    same_next = np.concatenate([sk[:-1] == sk[1:], [False]])
    multi = same_prev | same_next
    pf = np.zeros(N, np.uint32)
    pf[order] = np.arange(N, dtype=np.uint32) | (multi.astype(np.uint32) << 31)
    starts = np.flatnonzero(multi & ~same_prev)
    ends = np.flatnonzero(multi & ~same_next)
    n0 = np.minimum(ends - starts + 1, 32)
    return pf, np.stack([starts, sk[starts], n0, ends - starts + 1 >= 128], 1).astype(np.int32)   # last: "long run" flag


@pytest.mark.parametrize("n", [1, 5, 7, 8, 39, 511, 512, 513, 2500, 4096, 8192, 20000, 100003, 600001])
def test_finish_step_sums_in_aten_order(n):
    import fm_for_online_recommendation_b200 as pkg
    from oracle.deep import lib as olib
    lib = pkg.require_cuda()
    rng = np.random.RandomState(n)
    d = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    lv = rng.uniform(0, 3, n).astype(np.float32)
    dd, dl = torch.from_numpy(d).cuda(), torch.from_numpy(lv).cuda()
    bias = torch.tensor([0.5], device="cuda")
    loss = torch.zeros(1, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    assert lib.fmb_finish_step(p(dd), p(dl), n, p(bias), 0.0, 1, p(loss), None) == 0
    fp = C.POINTER(C.c_float)
    want_loss = np.float32(olib().orc_sum_aten(lv.ctypes.data_as(fp), n)) / np.float32(n)
    assert np.float32(loss.item()) == np.float32(want_loss)
    # SGD with lr=1 exposes the summed gradient: bias - 1*g
    bias.fill_(0.0)
    assert lib.fmb_finish_step(p(dd), None, n, p(bias), 1.0, 1, None, None) == 0
    want_g = np.float32(olib().orc_sum_aten(d.ctypes.data_as(fp), n))
    assert np.float32(bias.item()) == np.float32(0.0) - want_g
    out = torch.zeros(1, device="cuda")
    assert lib.fmb_sum_aten(p(dd), n, p(out), None) == 0
    assert np.float32(out.item()) == want_g


@pytest.mark.parametrize("sizes,k,B,real_xv,zipf,lr,scale", [
    ([943, 1682], 10, 256, False, False, 0.01, 1.0),      # cfg1 shape, raw N(0,1) init
    ([7, 5, 11, 3, 13, 4], 10, 700, True, True, 0.001, 0.2),  # heavy duplicates: long runs (> 512 entries/row)
    (FRAPPE, 64, 512, False, False, 0.001, 0.1),          # cfg3 shape
    (CRITEO, 10, 2500, False, False, 1e-4, 1.0),          # main_experiment.py pre-training shape
    ([5, 3], 4, 1, True, False, 0.01, 0.5),               # single sample
])
def test_update_embedding_and_fit_bit_exact(sizes, k, B, real_xv, zipf, lr, scale):
    m, orc = _pair("FMAdam", sizes, k, lr=lr, scale=scale)
    for step in range(4):
        Xi, Xv, Y = synth(sizes, B, 10 + step, real_xv=real_xv, zipf=zipf)
        if step % 2 == 0:
            got = float(m.update_embedding(Xi, Xv, Y).cpu())
            want = orc.update_embedding(Xi, Xv, Y)
            assert np.float32(got) == np.float32(want)
        else:
            m.fit(Xi, Xv, Y)
            orc.fit(Xi, Xv, Y)
        assert_same_params(m, orc, exact=True)


def test_sgd_mode_bit_exact():
    import fm_for_online_recommendation_b200 as pkg
    from oracle.deep import OracleDeep
    sizes = [50, 20, 7]
    orc = OracleDeep("FMAdam", sizes, 8, lr=0.05, update_mode=1, seed=3)
    orc.V *= np.float32(0.3)
    m = pkg.FMAdam(sizes, embedding_size=8, n=0.05, update_mode=1)
    push(m, orc)
    for step in range(3):
        Xi, Xv, Y = synth(sizes, 128, 40 + step, real_xv=True)
        m.update_embedding(Xi, Xv, Y)
        orc.update_embedding(Xi, Xv, Y)
    assert_same_params(m, orc, exact=True)


def test_host_entry_point_matches_device_entry_point():
    """fmb_session_fm_step_host (host buffers, copies inside) == the device-input step."""
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    sizes, k, B = CRITEO, 10, 2048
    m, orc = _pair("FMAdam", sizes, k, lr=1e-3, scale=0.2)
    Xi, Xv, Y = synth(sizes, B, 77)
    ids = orc.global_ids(Xi)
    y = np.ascontiguousarray(Y, dtype=np.float32)
    s = m._get_session(B)
    loss = C.c_float()
    rc = lib.fmb_session_fm_step_host(s, ids.ctypes.data_as(C.c_void_p), None, y.ctypes.data_as(C.c_void_p), B,
                                      C.c_void_p(m._table.data_ptr()), C.c_void_p(m.bias.data_ptr()), m._key_bits, 0,
                                      m._lr, 0, C.byref(loss), None)
    assert rc == 0, lib.fmb_last_error()
    want = orc.update_embedding(Xi, Xv, Y)
    assert np.float32(loss.value) == np.float32(want)
    assert_same_params(m, orc, exact=True)


def test_presorted_steps_bit_exact():
    """step(i); presort(i+1) -- the pre-sorted path must give the oracle's bits, also when a step arrives
    without (or with a stale) pre-sort."""
    sizes, k, B = [300, 40, 7, 2000, 3], 10, 700
    m, orc = _pair("FMAdam", sizes, k, lr=1e-3, scale=0.2)
    batches = [synth(sizes, B, 60 + i, zipf=(i % 2 == 1)) for i in range(7)]
    enc = [m.encode(*b) for b in batches]
    m.presort(enc[0])
    for i in range(7):
        loss = m._fm_step(enc[i], 0)
        if i in (0, 1, 3, 4):          # steps 3 and 6 run without a matching pre-sort
            m.presort(enc[i + 1])
        if i == 4:
            m.presort(enc[6])          # stale: batch 5 comes next
        want = orc.update_embedding(*batches[i])
        assert np.float32(loss.item()) == np.float32(want), i
        assert_same_params(m, orc, exact=True)


def test_large_batch_generic_sort_path_bit_exact():
    """B > 65 536: the step sorts with the generic multi-pass radix sort, the position words and the run list (one segment,
    short runs upwards / long runs downwards) come from pos_flags_kernel; the run kernel's long-run CTAs get 1 000-entry
    chains.  Losses and parameters == oracle, with and without the next batch's sort riding along."""
    sizes, k, B = [70, 9, 40000, 3000, 150000], 10, 70001
    m, orc = _pair("FMAdam", sizes, k, lr=1e-3, scale=0.2)
    batches = [synth(sizes, B, 80 + i, zipf=(i == 1)) for i in range(3)]
    enc = [m.encode(*b) for b in batches]
    for i in range(3):
        loss = m._fm_step(enc[i], 0, enc[i + 1] if i == 0 else None)
        want = orc.update_embedding(*batches[i])
        assert np.float32(loss.item()) == np.float32(want), i
    assert_same_params(m, orc, exact=True)


@pytest.mark.parametrize("nslot,nsteps,B", [(2, 5, 512), (4, 11, 512), (4, 9, 4100)])
def test_pipelined_host_entry_point_bit_exact(nslot, nsteps, B):
    """fmb_session_fm_step_host_async (nslot input slots, losses collected nslot - 1 steps later; pinned and pageable sources;
    pre-sort graphs and one step graph per slot at B = 4100) == oracle, step after step."""
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    assert lib.fmb_session_host_slots() >= nslot
    sizes, k = [300, 40, 7, 2000, 3], 10
    m, orc = _pair("FMAdam", sizes, k, lr=1e-3, scale=0.2)
    s = m._get_session(B)
    batches = [synth(sizes, B, 90 + i, zipf=(i % 2 == 0)) for i in range(nsteps)]
    ids = [orc.global_ids(b[0]) for b in batches]
    ys = [np.ascontiguousarray(b[2], dtype=np.float32) for b in batches]
    pinned = [torch.from_numpy(a).pin_memory() for a in ids]      # even steps: pinned, read in place
    tp, bp = C.c_void_p(m._table.data_ptr()), C.c_void_p(m.bias.data_ptr())
    loss = C.c_float()
    got = []
    lag = nslot - 1
    for i in range(nsteps):
        src = C.c_void_p(pinned[i].data_ptr()) if i % 2 == 0 else ids[i].ctypes.data_as(C.c_void_p)
        rc = lib.fmb_session_fm_step_host_async(s, i % nslot, src, None, ys[i].ctypes.data_as(C.c_void_p), B, tp, bp,
                                                m._key_bits, 0, m._lr, 0, None)
        assert rc == 0, lib.fmb_last_error()
        if i >= lag:
            assert lib.fmb_session_wait_loss(s, (i - lag) % nslot, C.byref(loss)) == 0
            got.append(loss.value)
    for i in range(max(0, nsteps - lag), nsteps):
        assert lib.fmb_session_wait_loss(s, i % nslot, C.byref(loss)) == 0
        got.append(loss.value)
    want = [orc.update_embedding(b[0], b[1], b[2]) for b in batches]
    assert [np.float32(g) for g in got] == [np.float32(w) for w in want]
    assert_same_params(m, orc, exact=True)


def test_full_size_properties_cfg4():
    """BASELINE cfg4 size (B=8192, Criteo tables): properties that need no oracle run.
    (1) rows not in the batch are untouched, (2) every touched row moves by <= lr per coordinate
    (fresh-Adam sign step), (3) the step is deterministic run to run."""
    import fm_for_online_recommendation_b200 as pkg
    sizes, k, B, lr = CRITEO, 10, 8192, 1e-4
    torch.manual_seed(0)
    m = pkg.FMAdam(sizes, embedding_size=k, n=lr)
    t0 = m._table.clone()
    b0 = m.bias.clone()
    Xi, Xv, Y = synth(sizes, B, 123)
    e = m.encode(Xi, Xv, Y)
    l1 = m.update_embedding(e, None, None).item()
    t1 = m._table.clone()
    touched = torch.zeros(m._R, dtype=torch.bool, device="cuda")
    touched[e.ids.reshape(-1).long()] = True
    assert torch.equal(t1[~touched], t0[~touched])
    # |delta| = lr up to the rounding of p - lr (parameters are O(1): one ulp is ~1.2e-7)
    assert float((t1 - t0).abs().max()) <= lr + 1e-6
    assert float((t1[touched][:, :k + 1] - t0[touched][:, :k + 1]).abs().max()) > 0
    with torch.no_grad():
        m._table.copy_(t0)
        m.bias.copy_(b0)
    l2 = m.update_embedding(e, None, None).item()
    assert l1 == l2 and torch.equal(m._table, t1)


def test_full_size_cfg5_steps_bit_exact_vs_oracle():
    """BASELINE cfg5 on one GPU, exactly as bench.py runs it: 39 fields, 33 M rows, k = 10, B = 8192, graph-replayed steps
    with the next batch's sort riding along (dense fields: cluster radix sort + tail; the 26 sparse fields: hash pass; run
    list; long runs first).  Losses and every touched row are bit-identical to the C oracle; untouched rows do not move."""
    import fm_for_online_recommendation_b200 as pkg
    from oracle.deep import OracleDeep
    sizes = [63, 113, 126, 51, 224, 148, 100, 79, 104, 9, 32, 57, 82] + [1_269_185] * 26
    k, B, lr = 10, 8192, 1e-4
    os.environ["ORC_THREADS"] = str(os.cpu_count() or 1)
    orc = OracleDeep("DeepFMAdam", sizes, k, 3, 400, lr=lr, seed=0)
    orc.w1 *= np.float32(0.3); orc.V *= np.float32(0.3)         # unsaturated logits: gradients of every magnitude
    m = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=lr)
    push(m, orc)
    t0 = m._table.clone()
    batches = [synth(sizes, B, 500 + i) for i in range(4)]
    enc = [m.encode(Xi, None, Y) for Xi, _, Y in batches]
    touched = torch.zeros(m._R, dtype=torch.bool, device="cuda")
    for i in range(4):
        got = float(m._fm_step(enc[i], 0, enc[i + 1] if i + 1 < 4 else None).cpu())
        want = orc.update_embedding(batches[i][0], batches[i][1], batches[i][2])
        assert np.float32(got) == np.float32(want), i
        touched[enc[i].ids.reshape(-1).long()] = True
    idx = torch.nonzero(touched).reshape(-1)
    rows = m._table[idx].cpu().numpy()
    ii = idx.cpu().numpy()
    assert np.array_equal(rows[:, :k], orc.V[ii]) and np.array_equal(rows[:, k], orc.w1[ii])
    assert np.float32(m.bias.item()) == np.float32(orc.bias.reshape(-1)[0])
    assert torch.equal(m._table[~touched], t0[~touched])


@pytest.mark.parametrize("op,fn", [(0, "orc_vec_sigmoid"), (1, "orc_vec_log_sigmoid"), (2, "orc_vec_sqrt_mkl"),
                                   (3, "orc_vec_expf_glibc")])
def test_aten_mirrors_device_equals_oracle(op, fn):
    """fmb_aten_math.cuh (device) == oracle_math.h (C) bit for bit: torch.sigmoid's Sleef/glibc split, log_sigmoid,
    MKL's vsSqrt (every mantissa), glibc expf.  The oracle side is pinned on torch itself in test_oracle_math.py."""
    import fm_for_online_recommendation_b200 as pkg
    from oracle.deep import lib as olib
    lib = pkg.require_cuda()
    rng = np.random.RandomState(op)
    if op == 2:
        xs = [((np.uint32(e0) << np.uint32(23)) + np.arange(1 << 24, dtype=np.uint32)).view(np.float32)
              for e0 in (1, 100, 127)]
        xs.append(np.arange(0, 1 << 23, dtype=np.uint32).view(np.float32))
    else:
        xs = [(rng.standard_normal(n) * rng.choice([0.05, 1, 5, 20, 60, 120], n)).astype(np.float32)
              for n in (1, 31, 33, 250, 8192, 8200, 1 << 22)]
        xs.append(np.array([0.0, -0.0, 88.72, 88.73, -103.9, -103.98, -103.3, -87.4, 1e-30, 104.5, -104.5],
                           np.float32))
    f = getattr(olib(), fn)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    for x in xs:
        x = np.ascontiguousarray(x)
        want = np.empty_like(x)
        f(x.ctypes.data, want.ctypes.data, x.size)
        d = torch.from_numpy(x).cuda()
        out = torch.empty_like(d)
        assert lib.fmb_math_eval(op, C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()), x.size, None) == 0
        got = out.cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (op, x.size)


@pytest.mark.parametrize("B", [512, 4096])
def test_step_graph_is_repointed_per_batch_and_next_batch_sort_rides_along(B):
    """fmb_session_fm_step_next: ONE graph per configuration, re-pointed at every call's batch; the sort of the
    announced next batch rides along.  12 distinct device-resident batches, bit-exact vs the oracle step by step;
    a wrong announcement (stepping on another batch than the one announced) must not change a bit either."""
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.require_cuda()
    m, orc = _pair("FMAdam", CRITEO, 10, lr=1e-3, scale=0.1)
    data = [synth(CRITEO, B, 100 + i, zipf=(i % 3 == 0)) for i in range(12)]
    enc = [m.encode(Xi, Xv, Y) for Xi, Xv, Y in data]
    order = [0, 1, 2, 3, 4, 5, 7, 6, 8, 9, 10, 11, 3, 3]      # step 6 was announced as batch 6 but batch 7 is stepped
    announce = [1, 2, 3, 4, 5, 6, 6, 8, 9, 10, 11, None, 3, None]
    for t, (i, nx) in enumerate(zip(order, announce)):
        got = m._fm_step(enc[i], 0, enc[nx] if nx is not None else None)
        want = orc.update_embedding(*data[i])
        assert np.float32(got.item()) == np.float32(want), (t, i)
    assert_same_params(m, orc)
    assert lib.fmb_session_graph_count(m._session) <= 6   # (sorted before / in the step) x (with / without next) x buffer parity


def test_ftrl_proximal_mode_bit_exact_and_sane():
    """update mode 2: per-coordinate FTRL-Proximal z/n/w fused into the step, bit-exact vs the oracle's
    restatement over 12 steps (duplicate-heavy and unique rows); the restatement itself is
    checked against a float64 evaluation of McMahan's closed form."""
    import ctypes as C
    from oracle.deep import lib as olib
    sizes, k, B = [9, 300, 7, 2000, 40], 6, 500
    m, orc = _pair("FMAdam", sizes, k, lr=0.05, scale=0.1)
    orc.enable_ftrl(1.0, 2e-3, 1e-3)
    m.enable_ftrl(1.0, 2e-3, 1e-3)
    for s in range(12):
        Xi, Xv, Y = synth(sizes, B, 900 + s, zipf=(s % 2 == 0))
        got = m.update_embedding(Xi, Xv, Y).item()
        want = orc.update_embedding(Xi, Xv, Y)
        assert np.float32(got) == np.float32(want), s
    assert_same_params(m, orc)
    zn = m._ftrl_zn.cpu().numpy()
    assert np.array_equal(zn[:, 0, :k], orc.fz_V) and np.array_equal(zn[:, 1, :k], orc.fn_V)
    assert np.array_equal(zn[:, 0, k], orc.fz_w1) and np.array_equal(zn[:, 1, k], orc.fn_w1)
    assert np.array_equal(m._ftrl_bias.cpu().numpy(), orc.f_bias)
