"""per-phase GPU time of the FUSED sharded step (csrc/shard3.cu) under torchrun: eager launches, CUDA events behind a sleep
kernel; the next batch's id exchange + owner sort are timed on the _pre stream."""
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
host = bench.synth_batches(sizes, B, 4, 1234 + rank)
enc = [m.encode(Xi, Y) for Xi, Y in host]
st = torch.cuda.current_stream()
names = ["issue_pre", "fused", "runs+finish", "join_pre"]
acc = np.zeros(len(names)); n = 0
pre = np.zeros(3)
m.prepare_fused(enc[0][0])
for it in range(30):
    y = enc[it % 4][1]; ids_next = enc[(it + 1) % 4][0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ep = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    dist.barrier(); torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    p = m._slot
    ev[0].record(st)
    pr = m._peer
    m._pre.wait_stream(st)
    with torch.cuda.stream(m._pre):
        ep[0].record(m._pre)
        check(m._lib.fmb_shard_transpose_ids_peers(ptr(ids_next), B, F, G, m.rank, pr["ptrs"][f"ids{1-p}"], *m._sync_args(), m.CH_IDS,
                                                   sh._stream()), "t")
        ep[1].record(m._pre)
        m._signal(m.CH_IDS, 2)
        ep[2].record(m._pre)
        m._sort_owned(pr["ids"][1 - p], 1 - p, pr["posflag"][1 - p])
        ep[3].record(m._pre)
    if os.environ.get("FMB_T3_SERIAL") == "1":
        st.wait_stream(m._pre)          # the fused kernel has the GPU to itself
    ev[1].record(st)
    ws, wsb = m._fused_launch(y, p, 0); ev[2].record(st)
    loss = m._fused_finish(ws, wsb, p); ev[3].record(st)
    st.wait_stream(m._pre); ev[4].record(st)
    m._slot = 1 - p
    torch.cuda.synchronize()
    if it >= 5:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(len(names))]); n += 1
        pre += np.array([ep[i].elapsed_time(ep[i + 1]) * 1000 for i in range(3)])
# per-tile phase stamps of one more step
import ctypes as C
T = m._peer["tiles"]
tdbg = torch.zeros(T * 8, dtype=torch.int64, device="cuda")
m._lib.fmb_debug_set_shard3_timestamps.argtypes = [C.c_void_p]; m._lib.fmb_debug_set_shard3_timestamps.restype = None
m._lib.fmb_debug_set_shard3_timestamps(C.c_void_p(tdbg.data_ptr()))
dist.barrier(); torch.cuda.synchronize()
p = m._slot
m._prepare_fused(enc[0][0], 1 - p); st.wait_stream(m._pre); torch.cuda.synchronize(); dist.barrier()
ws, wsb = m._fused_launch(enc[3][1], p, 0); m._fused_finish(ws, wsb, p); m._slot = 1 - p
torch.cuda.synchronize()
m._lib.fmb_debug_set_shard3_timestamps(None)
ts = tdbg.cpu().numpy().reshape(T, 8).astype(np.float64)
t0 = ts[:, 0].min()
ph = ["P0 ids+lists", "P1 gather", "P2+P3 partial+send", "P4 combine(own tiles)", "P5 wait ctx", "P6 update"]
if rank == 0:
    d = {ph[i]: round(float((ts[:, i + 1] - ts[:, i]).mean()) / 1e3, 2) for i in range(6)}
    own = (np.arange(T) % G) == rank
    d["P4 own tiles only"] = round(float((ts[own, 4] - ts[own, 3]).mean()) / 1e3, 2)
    d["tile start spread"] = round(float(ts[:, 0].max() - t0) / 1e3, 2)
    d["kernel span"] = round(float(ts[:, 6].max() - t0) / 1e3, 2)
    d["mean tile duration"] = round(float((ts[:, 6] - ts[:, 0]).mean()) / 1e3, 2)
    print(json.dumps({"tile_phase_us": d}), flush=True)
m.check_exchange(); m.check_overflow()
t = torch.tensor(np.concatenate([acc / n, pre / n]), device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    v = t.tolist()
    print(json.dumps({"world": world, "main_us": {nm: round(float(x), 1) for nm, x in zip(names, v[:4])},
                      "pre_us": {"transpose": round(v[4], 1), "wait_ids": round(v[5], 1), "owner_sort": round(v[6], 1)}}), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
