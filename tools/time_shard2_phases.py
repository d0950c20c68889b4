"""Per-phase device time of ShardedFM2.step (eager, CUDA events) under torchrun; prints rank 0's medians."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_for_online_recommendation_b200 import sharded2 as s2
from bench import feature_sizes, synth_batches
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
sizes = feature_sizes("cfg5"); B = 8192
m = s2.ShardedFM2(sizes, 10, B, n=1e-4, seed=0)
enc = [m.encode(Xi, Y) for Xi, Y in synth_batches(sizes, B, 4, 1234 + rank)]
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {}
m.phase_sort(enc[0][0], 0)
for t in range(14):
    p = t & 1
    e = [ev() for _ in range(5)]
    e[0].record(); m.phase_sort(enc[(t + 1) % 4][0], 1 - p)
    e[1].record(); m.phase_rows(p); m.phase_forward(enc[t % 4][0], enc[t % 4][1], p)
    e[2].record(); loss = m.phase_owner(p)
    e[3].record()
    torch.cuda.synchronize()
    if t >= 4:
        for n, a, b in (("sort+posflags+push+KEYS barrier", 0, 1), ("rows+ROWS barrier+fused+runs+PUSH barrier", 1, 2), ("owner+finish", 2, 3)):
            acc.setdefault(n, []).append(e[a].elapsed_time(e[b]) * 1e3)
# finer: the kernels of phase_forward / phase_owner one by one (barriers excluded from the kernel numbers)
import ctypes as C
from fm_for_online_recommendation_b200._lib import ptr
lib = m._lib
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
for t in range(14, 28):
    p = t & 1
    ids, y = enc[t % 4]
    m.phase_sort(enc[(t + 1) % 4][0], 1 - p)
    e = [ev() for _ in range(9)]
    e[7].record()
    m.phase_rows(p)
    e[0].record()
    lib.fmb_shard2_fused(ptr(ids), None, ptr(y), ptr(m.posflag[p]), m._peer_ptrs["table"], m._peer_ptrs["inbox"],
                         m._peer_ptrs[f"dl{p}"], ptr(m.rowbox), ptr(m.hot), ptr(m.hot_base_dev), ptr(m.field_off_dev),
                         ptr(m.bias), m.G, m.rank, m.B, m.F, m.k, 0, ptr(m.ws), m.ws_bytes, *m._wait(s2.CH_ROWS), st())
    e[1].record()
    lib.fmb_shard2_runs(ptr(m.skeys[p]), m.N, m.F, m.k, ptr(m.ws), m.ws_bytes, m._peer_ptrs["inbox"], m.G, m.rank, st())
    e[2].record()
    m._signal(s2.CH_PUSH, 1)
    e[3].record()
    lib.fmb_shard2_owner_apply(ptr(m.keys_all[p]), ptr(m.inbox), ptr(m.table), ptr(m.cnt), ptr(m.olist), ptr(m.nlist), p,
                               m._peer_ptrs["hot"], ptr(m.hot_base_dev), ptr(m.field_off_dev), m.G, m.rank, m.B, m.F, m.k,
                               m.lr, 0, *m._wait(s2.CH_PUSH), st())
    e[4].record()
    loss = torch.empty((), device="cuda")
    lib.fmb_finish_step(ptr(m.dl[p]), ptr(m.dl[p][m.G * m.B:]), m.G * m.B, ptr(m.bias), m.lr, 0, ptr(loss), st())
    e[5].record()
    e[6].record()
    torch.cuda.synchronize()
    if t >= 18:
        for n, a, b in (("  fused (waits ROWS; local gathers)", 0, 1), ("  runs (push partials)", 1, 2), ("  PUSH publish", 2, 3),
                        ("  owner (wait PUSH)+count+apply+reset", 3, 4), ("  finish", 4, 5), ("  push_rows + ROWS publish", 7, 0)):
            acc.setdefault(n, []).append(e[a].elapsed_time(e[b]) * 1e3)
dist.barrier()
if rank == 0:
    for n, v in acc.items():
        print(f"G={world} {n}: median {np.median(v):.1f} us")
os._exit(0)
