"""per-phase CUDA-event timing of the sharded step under torchrun (eager, main stream)."""
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
names = ["ids_T", "allgather_ids", "partial_fwd(+sort on side)", "all_to_all", "combine", "allgather_ctx", "backward+finish"]
acc = np.zeros(len(names)); n = 0
st = torch.cuda.current_stream()
for it in range(40):
    ids, y = enc[it % 4]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier(); torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)   # ~2 ms: the host enqueues the whole step meanwhile, so events see GPU time only
    ev[0].record(st)
    idsT = m.phase_ids(ids); ev[1].record(st)
    idsT_all = m._buf("idsT_all", (G, F, B), torch.int32)
    dist.all_gather_into_tensor(idsT_all.view(-1), idsT.view(-1)); ev[2].record(st)
    partial = m.phase_owner_forward(idsT_all); ev[3].record(st)
    recv = m._buf("recv", (G, B, m.PW))
    dist.all_to_all_single(recv.view(-1), partial.view(-1)); ev[4].record(st)
    ctx = m.phase_combine(recv, y); ev[5].record(st)
    ctx_all = m._buf("ctx_all", (G * B, m.CW))
    dist.all_gather_into_tensor(ctx_all.view(-1), ctx.view(-1)); ev[6].record(st)
    loss = m.phase_backward(ctx_all); ev[7].record(st)
    torch.cuda.synchronize()
    if it >= 10:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(len(names))]); n += 1
t = torch.tensor(acc / n, device="cuda")
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    out = {"world": world, "B_per_gpu": B, "phase_us_max_over_ranks": {nm: round(float(v), 1) for nm, v in zip(names, tmax.tolist())},
           "sum_us": round(float(tmax.sum()), 1), "bytes": {"allgather_ids_recv": G * F * B * 4, "all_to_all_send": G * B * m.PW * 4,
           "allgather_ctx_recv": G * B * m.CW * 4}}
    print(json.dumps(out), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
