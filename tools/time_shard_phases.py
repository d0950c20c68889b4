import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fm_for_online_recommendation_b200 import sharded as sh
G=int(sys.argv[1]) if len(sys.argv)>1 else 8
sizes=bench.feature_sizes('cfg5'); B=8192; F=len(sizes)
m=sh.ShardedFM(sizes, 10, n=1e-4, world=G, rank=0)
rng=np.random.RandomState(0)
off=m.offsets_np[:-1]
ids_all=[torch.from_numpy((np.stack([rng.randint(0,fs,size=B) for fs in sizes],1)+off[None,:]).astype(np.int32)).cuda() for r in range(G)]
y=torch.from_numpy((rng.uniform(size=B)<0.3).astype(np.float32)).cuda()
def ev(): return torch.cuda.Event(enable_timing=True)
names=['transpose','partial_fwd+sort','combine','backward+finish']
acc={n:0.0 for n in names}
reps=20
for it in range(reps+3):
    e=[ev() for _ in range(5)]
    e[0].record()
    idsT=[m.phase_ids(i).clone() for i in ids_all[:1]]
    e[1].record()
    idsT_all=torch.stack([m.phase_ids(i).clone() for i in ids_all]).contiguous() if it==0 else idsT_all
    torch.cuda.synchronize(); e[1].record()
    partial=m.phase_owner_forward(idsT_all)
    e[2].record()
    recv=partial.clone()  # stand-in for the all-to-all (same shape)
    ctx=m.phase_combine(recv,y)
    e[3].record()
    ctx_all=ctx.repeat(G,1).contiguous() if it==0 else ctx_all
    torch.cuda.synchronize(); e3b=ev(); e3b.record()
    loss=m.phase_backward(ctx_all)
    e[4].record(); torch.cuda.synchronize()
    if it>=3:
        acc['transpose']+=e[0].elapsed_time(e[1])/reps
        acc['partial_fwd+sort']+=e[1].elapsed_time(e[2])/reps
        acc['combine']+=e[2].elapsed_time(e[3])/reps
        acc['backward+finish']+=e3b.elapsed_time(e[4])/reps
print('G',G,{k:round(v*1000,1) for k,v in acc.items()},'us (includes python launch gaps)')
