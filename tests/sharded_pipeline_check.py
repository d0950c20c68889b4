"""Pipelined sharded step == un-pipelined sharded step, bit for bit (losses and every owned row), over NCCL.

Not collected by pytest (no test_ prefix): run as a program, one process per GPU,
    python tests/sharded_pipeline_check.py                      # world 1 (tests/test_sharded.py runs this)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/sharded_pipeline_check.py [--graph] [--peers]
--peers: the pipelined model exchanges through peer-mapped symmetric memory + epoch flags instead of NCCL.
--fused: the pipelined model runs the step as ONE kernel per rank with per-tile flags (csrc/shard3.cu).
Prints "PIPELINE_CHECK_OK world=<G> graph=<0|1>" on rank 0 and exits 0, or raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fm_for_online_recommendation_b200 import sharded as sh  # noqa: E402

SIZES = [7, 3, 40, 2, 1000, 13]


def main():
    graph = "--graph" in sys.argv
    peers = "--peers" in sys.argv
    fused = "--fused" in sys.argv
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    B, steps = 128, 7
    rng = np.random.RandomState(100 + rank)
    batches = []
    for _ in range(steps):
        Xi = np.stack([rng.randint(0, fs, size=B) for fs in SIZES], 1)
        Y = (rng.uniform(size=B) < 0.4).astype(np.float32)
        batches.append((Xi, Y))
    a = sh.ShardedFM(SIZES, 4, n=0.01, seed=3)
    b = sh.ShardedFM(SIZES, 4, n=0.01, seed=3)
    assert torch.equal(a.table, b.table)
    enc = [a.encode(Xi, Y) for Xi, Y in batches]
    if graph:
        # capture trains two warm-up steps on enc[0]: do the same on both models so they stay identical
        a.capture(*enc[0])
        (b.capture_fused if fused else b.capture_peers if peers else b.capture_pipelined)(*enc[0])
        assert torch.equal(a.table, b.table) and torch.equal(a.bias, b.bias)
        step_a, step_b = a.step_graphed, b.step_graphed_pipelined
    else:
        step_a, step_b = a.update_embedding, (b.update_embedding_fused if fused else b.update_embedding_peers if peers
                                              else b.update_embedding_pipelined)
    la = [float(step_a(ids, y).item()) for ids, y in enc]
    (b.prepare_fused if fused else b.prepare_peers if peers else b.prepare)(enc[0][0])
    lb = []
    for i in range(steps):
        nxt = enc[i + 1][0] if i + 1 < steps else (enc[0][0] if graph else None)
        lb.append(float(step_b(enc[i][1], nxt).item()))
    torch.cuda.synchronize()
    a.check_overflow()
    b.check_overflow()
    b.check_exchange()
    assert la == lb, (la, lb)
    assert torch.equal(a.table, b.table), "owned rows differ"
    assert torch.equal(a.bias, b.bias)
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print(f"PIPELINE_CHECK_OK world={world} graph={int(graph)} peers={int(peers)} fused={int(fused)} losses={la[:3]}", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)   # see sharded.bench_main: no communicator teardown after graph-captured collectives


if __name__ == "__main__":
    main()
