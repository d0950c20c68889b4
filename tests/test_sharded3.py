"""The sharded step as ONE kernel per rank (csrc/shard3.cu) with the ranks EMULATED in one process: every rank's kernel runs
on its own stream of the one GPU, the "peer" pointers are the other ranks' buffers in the same device memory, and the tiles
really wait for each other's flags (the tile counts are small enough for all G kernels to be resident together).  The result
must be bit-identical to the oracle in sharded order -- the same bar as the three-kernel path (test_sharded.py)."""
import numpy as np
import pytest
import torch

from fm_for_online_recommendation_b200 import sharded as sh
from test_sharded import K, SIZES, _oracle

pytestmark = pytest.mark.gpu


def _attach(ranks, B):
    G = len(ranks)
    arenas = []
    for m in ranks:
        _, total, T = m._peer_layout(B)
        assert T > 0
        arenas.append(torch.zeros(total, dtype=torch.int32, device=m.device))
    base = [a.data_ptr() for a in arenas]
    for m, a in zip(ranks, arenas):
        m._peer_bind(B, a, base)
    return arenas


@pytest.mark.parametrize("G,B,zipf_cap", [(1, 64, False), (2, 64, False), (4, 64, False), (8, 64, False), (2, 256, False), (4, 128, False)])
def test_fused_step_emulated_ranks_bit_exact(G, B, zipf_cap):
    steps = 3
    orc, (V0, w0, b0), batches, want_losses = _oracle(G, G * B, steps)
    ranks = [sh.ShardedFM(SIZES, K, n=0.01, init="zeros", world=G, rank=r) for r in range(G)]
    for m in ranks:
        m.load_full(V0, w0, b0)
    _attach(ranks, B)
    streams = [torch.cuda.Stream() for _ in range(G)]
    for (Xi, Y), want in zip(batches, want_losses):
        enc = [m.encode(Xi[r * B:(r + 1) * B], Y[r * B:(r + 1) * B]) for r, m in enumerate(ranks)]
        # ids exchange (the transpose kernel stores into every rank's slab; no flags needed in one process) + owner sort
        for m, e in zip(ranks, enc):
            pr = m._peer
            sh.check(m._lib.fmb_shard_transpose_ids_peers(sh.ptr(e[0]), B, m.F, G, m.rank, pr["ptrs"]["ids0"], *m._sync_args(), -1,
                                                          sh._stream()), "transpose")
        torch.cuda.synchronize()
        for m in ranks:
            m._sort_owned(m._peer["ids"][0], 0, m._peer["posflag"][0])
        torch.cuda.synchronize()
        # the fused kernels of all ranks, concurrently: they wait for each other's tile flags
        held = []
        for m, e, st in zip(ranks, enc, streams):
            with torch.cuda.stream(st):
                held.append(m._fused_launch(e[1], 0, 0))
        torch.cuda.synchronize()
        for m in ranks:
            m.check_exchange()
            m.check_overflow()
        losses = [float(m._fused_finish(ws, wsb, 0).item()) for m, (ws, wsb) in zip(ranks, held)]
        assert all(np.float32(l) == np.float32(want) for l in losses), (losses, want)
    for r, m in enumerate(ranks):
        V, w1 = m.local_params()
        assert np.array_equal(V, sh.shard_from_full(orc.V, G, r))
        assert np.array_equal(w1, sh.shard_from_full(orc.w1, G, r))
        assert np.float32(m.bias.item()) == np.float32(orc.bias[0])


@pytest.mark.parametrize("G,B", [(2, 256), (8, 128)])
def test_fused_step_hash_pass_of_sparse_fields_bit_exact(G, B):
    """fields with many more rows than owned entries take the owner sort's hash pass (only multi-hit entries are listed):
    two such fields (300 007 and 50 021 rows) whose ids are drawn from a few hundred distinct rows spread over the range, so
    that rows hit once, twice and many times all occur; still bit-identical to the oracle in sharded order"""
    from oracle.deep import OracleDeep
    sizes = [9, 300007, 5, 50021, 64, 3]
    k, steps = 6, 3
    orc = OracleDeep("FMAdam", sizes, k, lr=0.01, seed=1)
    orc.V *= np.float32(0.3)
    V0, w0, b0 = orc.V.copy(), orc.w1.copy(), orc.bias.copy()
    orc.set_shard_order(G)
    rs = np.random.RandomState(5)
    batches, want_losses = [], []
    for s in range(steps):
        n = G * B
        Xi = np.stack([rs.randint(0, 9, n), (rs.randint(0, 700, n) * 428) % 300007, rs.randint(0, 5, n),
                       (rs.randint(0, 2000, n) * 25) % 50021, rs.randint(0, 64, n), rs.randint(0, 3, n)], 1)
        Y = (rs.uniform(size=n) < 0.4).astype(np.float32)
        batches.append((Xi, Y))
        want_losses.append(orc.update_embedding(Xi, np.ones(Xi.shape, np.float32), Y))
    ranks = [sh.ShardedFM(sizes, k, n=0.01, init="zeros", world=G, rank=r) for r in range(G)]
    for m in ranks:
        m.load_full(V0, w0, b0)
    _attach(ranks, B)
    streams = [torch.cuda.Stream() for _ in range(G)]
    for (Xi, Y), want in zip(batches, want_losses):
        enc = [m.encode(Xi[r * B:(r + 1) * B], Y[r * B:(r + 1) * B]) for r, m in enumerate(ranks)]
        for m, e in zip(ranks, enc):
            pr = m._peer
            sh.check(m._lib.fmb_shard_transpose_ids_peers(sh.ptr(e[0]), B, m.F, G, m.rank, pr["ptrs"]["ids0"], *m._sync_args(), -1,
                                                          sh._stream()), "transpose")
        torch.cuda.synchronize()
        for m in ranks:
            m._sort_owned(m._peer["ids"][0], 0, m._peer["posflag"][0])
        torch.cuda.synchronize()
        # the hash pass lists only multi-hit entries: fewer than the owned entries of the big fields
        cnts = ranks[0]._ws["counts0"].cpu().numpy()
        assert cnts[1] < B and cnts[3] < B, cnts
        held = []
        for m, e, st in zip(ranks, enc, streams):
            with torch.cuda.stream(st):
                held.append(m._fused_launch(e[1], 0, 0))
        torch.cuda.synchronize()
        for m in ranks:
            m.check_exchange()
            m.check_overflow()
        losses = [float(m._fused_finish(ws, wsb, 0).item()) for m, (ws, wsb) in zip(ranks, held)]
        assert all(np.float32(l) == np.float32(want) for l in losses), (losses, want)
    for r, m in enumerate(ranks):
        V, w1 = m.local_params()
        assert np.array_equal(V, sh.shard_from_full(orc.V, G, r))
        assert np.array_equal(w1, sh.shard_from_full(orc.w1, G, r))


def test_fused_step_reports_tile_overflow_and_unsupported_shapes():
    m = sh.ShardedFM(SIZES, K, n=0.01, init="zeros", world=2, rank=0)
    assert m.fused_supported(64) and not m.fused_supported(40)          # B % (8 G)
    assert not sh.ShardedFM(SIZES, K, world=3, rank=0, init="zeros").fused_supported(48)   # G not a power of two
