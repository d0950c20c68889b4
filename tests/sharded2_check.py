"""ShardedFM2 (one process per GPU, peer memory + epoch flags) == the oracle's rank-partial step on the concatenated
batch, bit for bit: losses of every step and every owned row.

Not collected by pytest (no test_ prefix): run as a program, one process per GPU,
    python tests/sharded2_check.py [--graph]                                   # world 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/sharded2_check.py [--graph]
Prints "SHARD2_CHECK_OK world=<G> graph=<0|1>" on rank 0 and exits 0, or raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fm_for_online_recommendation_b200 import sharded2 as s2  # noqa: E402
from oracle.deep import OracleDeep  # noqa: E402

SIZES = [7, 3, 40, 2, 1000, 13, 5000, 64]
K = 6


def main():
    graph = "--graph" in sys.argv
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    B, steps = 256, 9
    rng = np.random.RandomState(7)
    batches = []
    for s in range(steps + 1):
        if s % 3 == 2:
            Xi = np.stack([np.minimum(rng.zipf(1.3, size=world * B) - 1, fs - 1) for fs in SIZES], 1)
        else:
            Xi = np.stack([rng.randint(0, fs, size=world * B) for fs in SIZES], 1)
        Y = (rng.uniform(size=world * B) < 0.4).astype(np.float32)
        batches.append((Xi, Y))
    orc = OracleDeep("FMAdam", SIZES, K, lr=0.01, seed=3)
    orc.V *= np.float32(0.3)
    orc.set_rank_partial_order(B)
    m = s2.ShardedFM2(SIZES, K, B, n=0.01, init="zeros", hot_max=100)
    m.load_full(orc.V, orc.w1, orc.bias)
    m.sync_hot()
    torch.cuda.synchronize()
    dist.barrier()
    mine = lambda s: m.encode(batches[s][0][rank * B:(rank + 1) * B], batches[s][1][rank * B:(rank + 1) * B])
    enc = [mine(s) for s in range(steps + 1)]
    if graph:
        # capture trains two warm-up steps on batch 0: mirror them in the oracle
        m.capture(*enc[0])
        for _ in range(2):
            orc.update_embedding(batches[0][0], np.ones(batches[0][0].shape, np.float32), batches[0][1])
        m.prepare(enc[0][0])
        m._g_ids[0].copy_(enc[0][0])
    else:
        m.prepare(enc[0][0])
    for s in range(steps):
        if graph:
            loss = m.step_graphed(enc[s][1], enc[s + 1][0])
        else:
            loss = m.step(enc[s][0], enc[s][1], enc[s + 1][0])
        want = orc.update_embedding(batches[s][0], np.ones(batches[s][0].shape, np.float32), batches[s][1])
        got = float(loss.item())
        assert np.float32(got) == np.float32(want), (rank, s, got, want)
    torch.cuda.synchronize()
    m.check_exchange()
    V, w1 = m.local_params()
    assert np.array_equal(V, s2.shard_from_full(orc.V, world, rank)), rank
    assert np.array_equal(w1, s2.shard_from_full(orc.w1, world, rank)), rank
    assert np.float32(m.bias.item()) == np.float32(orc.bias[0])
    dist.barrier()
    if rank == 0:
        print(f"SHARD2_CHECK_OK world={world} graph={int(graph)}", flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
