import sys, os, ctypes as C, numpy as np, torch
os.environ['FMB_NO_GRAPH']='1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fm_for_online_recommendation_b200 as pkg
lib=pkg.require_cuda()
sizes=bench.feature_sizes(os.environ.get('WL','cfg5')); B=8192
torch.manual_seed(0)
m=pkg.FMAdam(sizes, embedding_size=10, n=1e-4) if os.environ.get('MODEL','fm')=='fm' else pkg.DeepFMAdam(sizes, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
host=bench.synth_batches(sizes,B,3,1)
enc=[m.encode(Xi,None,Y) for Xi,Y in host]
step=(lambda e: m._fm_step(e,0)) if os.environ.get('MODEL','fm')=='fm' else (lambda e: m._deep_fit(e))
for i in range(3): step(enc[i%3])
torch.cuda.synchronize()
dbg=torch.zeros(8+4000*8,dtype=torch.int64,device='cuda')
lib.fmb_debug_set_runs_buffer.argtypes=[C.c_void_p]; lib.fmb_debug_set_runs_buffer.restype=None
dbg[1]=int(os.environ.get('MINLEN','0'))
lib.fmb_debug_set_runs_buffer(C.c_void_p(dbg.data_ptr()))
step(enc[0]); torch.cuda.synchronize()
lib.fmb_debug_set_runs_buffer(None)
d=dbg.cpu().numpy(); n=int(d[0]); r=d[8:8+min(n,4000)*8].reshape(-1,8)
print('runs',n)
t0=r[:,1].min()
end=r[:,1]-t0+r[:,2]+r[:,3]+r[:,4]
order=np.argsort(-end)
print('len start direct ring update c_issue c_wait c_cons')
for i in order[:12]: print(r[i,0], r[i,1]-t0, r[i,2], r[i,3], r[i,4], r[i,5], r[i,6], r[i,7])
print('by length buckets: lo hi count mean_direct mean_ring mean_update | issue wait cons')
for lo,hi in [(2,8),(8,32),(32,64),(64,128),(128,256),(256,512),(512,2000),(2000,100000)]:
    s=(r[:,0]>=lo)&(r[:,0]<hi)
    if s.any(): print(lo,hi,int(s.sum()), r[s,2].mean().round(), r[s,3].mean().round(), r[s,4].mean().round(), '|', r[s,5].mean().round(), r[s,6].mean().round(), r[s,7].mean().round())
