"""Emulated ranks (one process, one GPU) at the bench's cfg5 shape: for ncu launch lists of the shard2 kernels."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_for_online_recommendation_b200 import sharded2 as s2
from bench import feature_sizes, synth_batches
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sizes = feature_sizes("cfg5"); B = 8192
ranks = [s2.ShardedFM2(sizes, 10, B, n=1e-4, seed=0, world=G, rank=r) for r in range(G)]
s2.ShardedFM2.bind_emulated(ranks)
for m in ranks: m.sync_hot()
torch.cuda.synchronize()
enc = [[m.encode(Xi, Y) for Xi, Y in synth_batches(sizes, B, 2, 1234 + r)] for r, m in enumerate(ranks)]
for t in range(3):
    slot = t & 1
    for r, m in enumerate(ranks): m.phase_sort(enc[r][t % 2][0], slot)
    torch.cuda.synchronize()
    for m in ranks: m.phase_rows(slot)
    torch.cuda.synchronize()
    for r, m in enumerate(ranks): m.phase_forward(enc[r][t % 2][0], enc[r][t % 2][1], slot)
    torch.cuda.synchronize()
    for m in ranks: m.phase_owner(slot)
    torch.cuda.synchronize()
print("ok")
