"""per-phase GPU time of the peer-memory sharded step (events on main behind a sleep kernel), fused vs separate signals."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fm_for_online_recommendation_b200 import sharded as sh
from fm_for_online_recommendation_b200._lib import check, ptr
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = bench.feature_sizes("cfg5"); B = 8192; k = 10
m = sh.ShardedFM(sizes, k, n=1e-4, seed=0)
G, F = m.G, m.F
lib = m._lib
host = bench.synth_batches(sizes, B, 4, 1234 + rank)
enc = [m.encode(Xi, Y) for Xi, Y in host]
st = torch.cuda.current_stream()
S = lambda: sh._stream()
out = {}
for fused in (1, 0):
    names = ["prepare_issue", "partial(+publish)", "combine(+wait,+bcast,+publish)", "wait_ctx", "backward+finish", "join_pre"]
    acc = np.zeros(len(names)); n = 0
    m.prepare_peers(enc[0][0])
    pr = m._peer
    for it in range(35):
        y = enc[it % 4][1]; ids_next = enc[(it + 1) % 4][0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier(); torch.cuda.synchronize()
        torch.cuda._sleep(4_000_000)
        p = m._slot
        ev[0].record(st)
        m._prepare_peers(ids_next, 1 - p); ev[1].record(st)
        if fused:
            check(lib.fmb_shard_partial_forward_peers(ptr(pr["ids"][p]), ptr(m.table), G, m.rank, B, F, k, pr["ptrs"]["recv"], *m._sync_args(), m.CH_A2A, S()), "pf"); ev[2].record(st)
            check(lib.fmb_shard_combine_peers(ptr(pr["recv"]), ptr(m.bias), ptr(y), G, m.rank, B, k, 0, pr["ptrs"]["ctx"], None, *m._sync_args(), m.CH_A2A, m.CH_CTX, S()), "cb"); ev[3].record(st)
        else:
            check(lib.fmb_shard_partial_forward_peers(ptr(pr["ids"][p]), ptr(m.table), G, m.rank, B, F, k, pr["ptrs"]["recv"], *m._sync_args(), -1, S()), "pf")
            m._signal(m.CH_A2A, 3); ev[2].record(st)
            ctx = m.phase_combine(pr["recv"], y, 0)
            check(lib.fmb_shard_ctx_bcast_peers(ptr(ctx), G, m.rank, B, k, pr["ptrs"]["ctx"], S()), "bc")
            m._signal(m.CH_CTX, 1); ev[3].record(st)
        m._signal(m.CH_CTX, 2); ev[4].record(st)
        loss = m.phase_backward(pr["ctx"], p, join_sort=False); ev[5].record(st)
        st.wait_stream(m._pre); ev[6].record(st)
        m._slot = 1 - p
        torch.cuda.synchronize()
        if it >= 5:
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(len(names))]); n += 1
    t = torch.tensor(acc / n, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["fused" if fused else "separate"] = dict({nm: round(float(v), 1) for nm, v in zip(names, t.tolist())}, sum=round(float(t.sum()), 1))
m.check_exchange()
if rank == 0: print(json.dumps({"world": world, "us": out}), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
