"""per-phase GPU time of the PIPELINED sharded step (events on main behind a sleep kernel)."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fm_for_online_recommendation_b200 import sharded as sh
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = bench.feature_sizes("cfg5"); B = 8192; k = 10
m = sh.ShardedFM(sizes, k, n=1e-4, seed=0)
G, F = m.G, m.F
host = bench.synth_batches(sizes, B, 4, 1234 + rank)
enc = [m.encode(Xi, Y) for Xi, Y in host]
st = torch.cuda.current_stream()
names = ["prepare_issue(ids_T)", "partial_fwd", "all_to_all", "combine", "allgather_ctx", "backward+finish", "join_pre"]
acc = np.zeros(len(names)); n = 0
sort_us = 0.0
m.prepare(enc[0][0])
for it in range(35):
    y = enc[it % 4][1]; ids_next = enc[(it + 1) % 4][0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    p = m._slot
    ev[0].record(st)
    # --- _prepare inlined to time the sort on _pre
    idsT = m.phase_ids(ids_next, 1 - p)
    idsT_all_n = m._buf(f"idsT_all{1-p}", (G, F, B), torch.int32)
    work = dist.all_gather_into_tensor(idsT_all_n.view(-1), idsT.view(-1), async_op=True)
    m._pre.wait_stream(st)
    with torch.cuda.stream(m._pre):
        work.wait(); es0.record(m._pre); m._sort_owned(idsT_all_n, 1 - p); es1.record(m._pre)
    ev[1].record(st)
    idsT_all = m._ws[f"idsT_all{p}"]
    partial = m.phase_partial(idsT_all.view(G, F, B)); ev[2].record(st)
    recv = m._buf("recv", (G, B, m.PW))
    dist.all_to_all_single(recv.view(-1), partial.view(-1)); ev[3].record(st)
    ctx = m.phase_combine(recv, y); ev[4].record(st)
    ctx_all = m._buf("ctx_all", (G * B, m.CW))
    dist.all_gather_into_tensor(ctx_all.view(-1), ctx.view(-1)); ev[5].record(st)
    loss = m.phase_backward(ctx_all, p, join_sort=False); ev[6].record(st)
    st.wait_stream(m._pre); ev[7].record(st)
    m._slot = 1 - p
    torch.cuda.synchronize()
    if it >= 5:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(len(names))]); n += 1
        sort_us += es0.elapsed_time(es1) * 1000
        first = ev[0].elapsed_time(es0) * 1000
t = torch.tensor(acc / n, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "phase_us": {nm: round(float(v), 1) for nm, v in zip(names, t.tolist())}, "sum": round(float(t.sum()), 1),
                      "sort_on_pre_us": round(sort_us / n, 1), "sort_start_after_us(last it)": round(first, 1)}), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
