"""Second multi-GPU design (csrc/shard2.cu, sharded2.py): rows gathered from the owners over peer memory, per-rank
partial gradients into the owners' inboxes, rank-ordered add at the owner.

GPU tests emulate the G ranks inside one process on one device (every rank's buffers live on the same GPU, the "peer"
pointers are the other ranks' buffers, the phases run in lock step and the epoch flags are skipped -- kernels that
spin on one another must never share a GPU).  The oracle restates the summation order (rank_B) and is itself compared
with the reference order.  tests/sharded2_check.py runs the real thing, one process per GPU, under torchrun.
"""
import os
import socket

import numpy as np
import pytest

from _util import synth

SIZES = [9, 40, 7, 3, 100, 23, 2, 64, 3000, 517]
K = 6


def _oracle(G, B, steps, rank_partial=True, seed=0, lr=0.01):
    from oracle.deep import OracleDeep
    orc = OracleDeep("FMAdam", SIZES, K, lr=lr, seed=seed)
    orc.V *= np.float32(0.3)
    V0, w0, b0 = orc.V.copy(), orc.w1.copy(), orc.bias.copy()
    if rank_partial:
        orc.set_rank_partial_order(B)
    losses, batches = [], []
    for s in range(steps):
        Xi, _, Y = synth(SIZES, G * B, 50 + s, zipf=(s % 2 == 1))
        batches.append((Xi, Y))
        losses.append(orc.update_embedding(Xi, np.ones(Xi.shape, np.float32), Y))
    return orc, (V0, w0, b0), batches, losses


def test_rank_partial_order_is_the_reference_order_for_one_rank_and_close_to_it_otherwise():
    """oracle only: rank_B = whole batch reproduces the reference order bit for bit; with 8 ranks the tables stay within
    1e-5 of the reference-order run after 40 steps (the distance SURVEY.md 8e asks to be measured)."""
    a, _, _, la = _oracle(1, 512, 6, rank_partial=False)
    b, _, _, lb = _oracle(1, 512, 6, rank_partial=True)
    assert np.array_equal(a.V, b.V) and np.array_equal(a.w1, b.w1) and la == lb
    ref, _, _, lr_ = _oracle(8, 64, 40, rank_partial=False, lr=1e-3)
    rp, _, _, lp = _oracle(8, 64, 40, rank_partial=True, lr=1e-3)
    frac_equal = float((ref.V == rp.V).mean())
    err = float(np.max(np.abs(ref.V - rp.V) / np.maximum(np.abs(ref.V), 1e-3)))
    assert frac_equal > 0.5 and err < 5e-3, (frac_equal, err)   # sign-step flips of 2*lr on near-zero gradients
    np.testing.assert_allclose(lp, lr_, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("G,B,hot_max", [(1, 64, 4096), (2, 48, 50), (4, 33, 0), (8, 16, 50), (3, 20, 4096), (8, 128, 50),
                                          (8, 256, 0)])
def test_emulated_ranks_bit_exact_vs_oracle(G, B, hot_max):
    import torch
    from fm_for_online_recommendation_b200 import sharded2 as s2
    steps = 4
    orc, (V0, w0, b0), batches, want = _oracle(G, B, steps)
    ranks = [s2.ShardedFM2(SIZES, K, B, n=0.01, init="zeros", world=G, rank=r, hot_max=hot_max) for r in range(G)]
    s2.ShardedFM2.bind_emulated(ranks)
    for m in ranks:
        m.load_full(V0, w0, b0)
    for m in ranks:
        m.sync_hot()
    torch.cuda.synchronize()
    for t, ((Xi, Y), w) in enumerate(zip(batches, want)):
        slot = t & 1
        enc = [m.encode(Xi[r * B:(r + 1) * B], Y[r * B:(r + 1) * B]) for r, m in enumerate(ranks)]
        for m, e in zip(ranks, enc):
            m.phase_sort(e[0], slot)
        torch.cuda.synchronize()
        for m in ranks:
            m.phase_rows(slot)
        torch.cuda.synchronize()
        for m, e in zip(ranks, enc):
            m.phase_forward(e[0], e[1], slot)
        torch.cuda.synchronize()
        losses = [float(m.phase_owner(slot).item()) for m in ranks]
        torch.cuda.synchronize()
        assert all(np.float32(l) == np.float32(w) for l in losses), (t, losses, w)
    for r, m in enumerate(ranks):
        V, w1 = m.local_params()
        assert np.array_equal(V, s2.shard_from_full(orc.V, G, r)), r
        assert np.array_equal(w1, s2.shard_from_full(orc.w1, G, r)), r
        assert np.float32(m.bias.item()) == np.float32(orc.bias[0])
        assert int(m.cnt.abs().sum().item()) == 0          # the owner's counters are back to zero


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_world1_process_group_step_equals_single_gpu_step(graph):
    """ShardedFM2 over a real process group (world 1 here: symmetric memory, flags, side stream, graphs) against the
    oracle; tests/sharded2_check.py under torchrun covers world > 1."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(here, "sharded2_check.py")] + (["--graph"] if graph else [])
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHARD2_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
