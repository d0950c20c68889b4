"""Row-sharded multi-GPU step.

CPU (gloo, world_size 2): the decomposition itself -- shard map, block layout of the all-gather /
all-to-all, owner-major reduction order -- is executed with numpy kernels over real
torch.distributed collectives and compared bit for bit with the oracle in sharded order.
GPU: the CUDA kernels of every rank are run in ONE process (ranks emulated, exchanges done by hand,
as B200_PROFILING.md prescribes when there are fewer GPUs than ranks) and compared bit for bit with
the same oracle.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import synth
from fm_for_online_recommendation_b200 import sharded as sh

SIZES = [9, 40, 7, 3, 100, 23, 2, 64]
K = 6


def _oracle(G, B_total, steps, seed=0, lr=0.01):
    from oracle.deep import OracleDeep
    orc = OracleDeep("FMAdam", SIZES, K, lr=lr, seed=seed)
    orc.V *= np.float32(0.3)
    V0, w0, b0 = orc.V.copy(), orc.w1.copy(), orc.bias.copy()
    orc.set_shard_order(G)
    losses = []
    batches = []
    for s in range(steps):
        Xi, _, Y = synth(SIZES, B_total, 50 + s, zipf=(s % 2 == 1))
        batches.append((Xi, Y))
        losses.append(orc.update_embedding(Xi, np.ones(Xi.shape, np.float32), Y))
    return orc, (V0, w0, b0), batches, losses


def test_shard_map_roundtrip():
    R = 37
    full = np.arange(R * 3, dtype=np.float32).reshape(R, 3)
    for G in (1, 2, 3, 8):
        shards = [sh.shard_from_full(full, G, r) for r in range(G)]
        assert [s.shape[0] for s in shards] == [sh.local_rows_count(R, G, r) for r in range(G)]
        assert np.array_equal(sh.full_from_shards(shards), full)
        for gid in range(R):
            assert shards[sh.owner_of(gid, G)][sh.local_row(gid, G), 0] == full[gid, 0]


# ---------------------------------------------------------------------------------------------- CPU / gloo
def _adam1(p, g, lr):
    import ctypes as C
    from oracle.deep import lib
    p = np.ascontiguousarray(p, np.float32).copy()
    g = np.ascontiguousarray(g, np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_update_dense(p.ctypes.data_as(fp), g.ctypes.data_as(fp), p.size, C.c_float(np.float32(lr)), 0)
    return p


def _cpu_rank(rank, world, port, B, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from oracle.deep import lib
    f32 = np.float32
    orc, (V0, w0, b0), batches, want_losses = _oracle(world, world * B, steps)
    V = sh.shard_from_full(V0, world, rank).copy()
    w1 = sh.shard_from_full(w0, world, rank).copy()
    bias = b0.copy()
    off = np.concatenate([[0], np.cumsum(SIZES)])[:-1]
    F, k = len(SIZES), K
    lr = f32(0.01)
    for (Xi, Y), want in zip(batches, want_losses):
        ids_local = (Xi[rank * B:(rank + 1) * B] + off[None, :]).astype(np.int32)
        y_local = Y[rank * B:(rank + 1) * B].astype(f32)
        # 1. all-gather ids (transposed, like the GPU path)
        idsT = torch.from_numpy(np.ascontiguousarray(ids_local.T))
        gathered = torch.empty(world, F, B, dtype=torch.int32)
        dist.all_gather_into_tensor(gathered.view(-1), idsT.view(-1))
        ids_all = gathered.numpy().transpose(0, 2, 1).reshape(world * B, F)
        # 2. owner partials for all samples: [S | Q | first], field order
        part = np.zeros((world * B, 2 * k + 1), f32)
        for bg in range(world * B):
            S = np.zeros(k, f32); Q = np.zeros(k, f32); fs = f32(0)
            for f in range(F):
                g = int(ids_all[bg, f])
                if sh.owner_of(g, world) != rank:
                    continue
                e = V[sh.local_row(g, world)]
                S = (S + e).astype(f32); Q = (Q + (e * e).astype(f32)).astype(f32)
                fs = f32(fs + w1[sh.local_row(g, world)])
            part[bg, :k] = S; part[bg, k:2 * k] = Q; part[bg, 2 * k] = fs
        # 3. all-to-all: block r of `part` goes to rank r
        recv = torch.empty(world, B, 2 * k + 1)
        dist.all_to_all_single(recv.view(-1), torch.from_numpy(part).view(-1))
        recv = recv.numpy()
        # 4. combine in owner order, logit, loss, delta (mean over the GLOBAL batch)
        S = np.zeros((B, k), f32); Q = np.zeros((B, k), f32); fs = np.zeros(B, f32)
        for o in range(world):
            S = (S + recv[o, :, :k]).astype(f32); Q = (Q + recv[o, :, k:2 * k]).astype(f32)
            fs = (fs + recv[o, :, 2 * k]).astype(f32)
        bi = (((S * S).astype(f32) - Q).astype(f32) * f32(0.5)).astype(f32)
        fp = C.POINTER(C.c_float)
        sb = np.array([lib().orc_sum_aten(np.ascontiguousarray(bi[b]).ctypes.data_as(fp), k) for b in range(B)], f32)
        z = ((fs + sb).astype(f32) + bias[0]).astype(f32)
        delta = np.empty(B, f32)
        # orc_loss_delta divides by its own B: feed the global batch size through a padded call
        zg = torch.empty(world, B); dist.all_gather_into_tensor(zg.view(-1), torch.from_numpy(z))
        yg = torch.empty(world, B); dist.all_gather_into_tensor(yg.view(-1), torch.from_numpy(y_local))
        zall = zg.numpy().reshape(-1).copy(); yall = yg.numpy().reshape(-1).copy()
        dall = np.empty(world * B, f32)
        loss = lib().orc_loss_delta(0, zall.ctypes.data_as(fp), yall.ctypes.data_as(fp), world * B,
                                    dall.ctypes.data_as(fp))
        assert f32(loss) == f32(want)
        # 5. all-gather context (S, delta)
        Sg = torch.empty(world, B, k); dist.all_gather_into_tensor(Sg.view(-1), torch.from_numpy(S).view(-1))
        S_all = Sg.numpy().reshape(world * B, k)
        # 6. owned rows: sum duplicates in global sample order, one fresh-Adam step per row
        gV = {}; gw = {}
        for bg in range(world * B):
            for f in range(F):
                g = int(ids_all[bg, f])
                if sh.owner_of(g, world) != rank:
                    continue
                lr_ = sh.local_row(g, world)
                d = dall[bg]
                e = V[lr_]
                ge = ((d * S_all[bg]).astype(f32) - (d * e).astype(f32)).astype(f32)
                gV[lr_] = (gV.get(lr_, np.zeros(k, f32)) + ge).astype(f32)
                gw[lr_] = f32(gw.get(lr_, f32(0)) + d)
        for lr_ in gV:
            V[lr_] = _adam1(V[lr_], gV[lr_], lr)
            w1[lr_] = _adam1(np.array([w1[lr_]], f32), np.array([gw[lr_]], f32), lr)[0]
        gb = f32(lib().orc_sum_aten(dall.ctypes.data_as(fp), world * B))
        bias = _adam1(bias, np.array([gb], f32), lr)
    # compare this rank's shard with the oracle's full table
    ok = (np.array_equal(V, sh.shard_from_full(orc.V, world, rank)) and
          np.array_equal(w1, sh.shard_from_full(orc.w1, world, rank)) and np.array_equal(bias, orc.bias))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_decomposition_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cpu_rank, args=(2, port, 12, 3, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("G,B", [(1, 64), (2, 48), (4, 33), (8, 16), (3, 20), (2, 64), (8, 128), (3, 64)])  # B % 64 == 0: warp-per-sample partial forward
def test_cuda_ranks_emulated_bit_exact(G, B):
    steps = 3
    orc, (V0, w0, b0), batches, want_losses = _oracle(G, G * B, steps)
    ranks = [sh.ShardedFM(SIZES, K, n=0.01, init="zeros", world=G, rank=r) for r in range(G)]
    for m in ranks:
        m.load_full(V0, w0, b0)
    for (Xi, Y), want in zip(batches, want_losses):
        enc = [m.encode(Xi[r * B:(r + 1) * B], Y[r * B:(r + 1) * B]) for r, m in enumerate(ranks)]
        idsT_all = torch.stack([m.phase_ids(e[0]).clone() for m, e in zip(ranks, enc)]).contiguous()
        partials = [m.phase_owner_forward(idsT_all).clone() for m in ranks]          # [G,B,PW] each
        ctxs = []
        for r, m in enumerate(ranks):
            recv = torch.stack([partials[o][r] for o in range(G)]).contiguous()       # the all-to-all
            ctxs.append(m.phase_combine(recv, enc[r][1]).clone())
        ctx_all = torch.cat(ctxs).contiguous()                                        # the all-gather
        losses = [float(m.phase_backward(ctx_all).item()) for m in ranks]
        assert all(np.float32(l) == np.float32(want) for l in losses)
    for r, m in enumerate(ranks):
        m.check_overflow()
        V, w1 = m.local_params()
        assert np.array_equal(V, sh.shard_from_full(orc.V, G, r))
        assert np.array_equal(w1, sh.shard_from_full(orc.w1, G, r))
        assert np.float32(m.bias.item()) == np.float32(orc.bias[0])


@pytest.mark.gpu
@pytest.mark.parametrize("graph,peers", [(False, False), (True, False), (False, True), (True, True), (False, "fused"),
                                         (True, "fused")])
def test_pipelined_step_equals_unpipelined_world1_nccl(graph, peers):
    """ShardedFM.update_embedding_pipelined (ids of batch t+1 exchanged and sorted under step t) against
    update_embedding over a real NCCL group (world 1 here; tests/sharded_pipeline_check.py under torchrun for
    world > 1): identical losses and tables.  Runs as a child process with a timeout."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    cmd = ([sys.executable, os.path.join(here, "sharded_pipeline_check.py")] + (["--graph"] if graph else []) +
           (["--fused"] if peers == "fused" else ["--peers"] if peers else []))   # --fused: one kernel per rank; --peers: exchanges through peer-mapped symmetric memory + epoch flags
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PIPELINE_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
