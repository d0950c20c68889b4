import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fm_for_online_recommendation_b200 import sharded as sh
from fm_for_online_recommendation_b200._lib import ptr
for G in (8, 2):
    sizes = bench.feature_sizes('cfg5'); B = 8192; F = len(sizes)
    m = sh.ShardedFM(sizes, 10, n=1e-4, world=G, rank=0)
    rng = np.random.RandomState(0); off = m.offsets_np[:-1]
    ids_all = [torch.from_numpy((np.stack([rng.randint(0, fs, size=B) for fs in sizes], 1) + off[None, :]).astype(np.int32)).cuda() for r in range(G)]
    idsT_all = torch.stack([m.phase_ids(i).clone() for i in ids_all]).contiguous()
    flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.phase_partial(idsT_all); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    print('G', G, 'partial forward us (L2 flushed): median', round(float(np.median(ts[2:])), 1), 'min', round(min(ts[2:]), 1))
    del m
