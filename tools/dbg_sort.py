import sys, os, ctypes as C, numpy as np, torch
os.environ['FMB_NO_GRAPH']='1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fm_for_online_recommendation_b200 as pkg
lib=pkg.require_cuda()
sizes=bench.feature_sizes('cfg5'); B=8192
torch.manual_seed(0)
m=pkg.FMAdam(sizes, embedding_size=10, n=1e-4)
host=bench.synth_batches(sizes,B,3,1)
enc=[m.encode(Xi,None,Y) for Xi,Y in host]
for i in range(3): m._fm_step(enc[i%3],0)
torch.cuda.synchronize()
dbg=torch.zeros(16,dtype=torch.int64,device='cuda')
lib.fmb_debug_set_sort_buffer.argtypes=[C.c_void_p]; lib.fmb_debug_set_sort_buffer.restype=None
lib.fmb_debug_set_sort_buffer(C.c_void_p(dbg.data_ptr()))
m._fm_step(enc[0],0); torch.cuda.synchronize()
lib.fmb_debug_set_sort_buffer(None)
names=['load','zero+rank','prefix','stage scan+scatter','cluster.sync#1','base calc','remote copy','cluster.sync#2','final store']
d=dbg.cpu().numpy()
for n_,v in zip(names,d): print(f'{n_:22s} {v}')
print('total', d[:9].sum())
