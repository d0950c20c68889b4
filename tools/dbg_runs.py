# per-run timing of the run kernel inside one eager FM step (run-list mode): start = ns since the first run started
import sys, os, ctypes as C, numpy as np, torch
os.environ['FMB_NO_GRAPH'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fm_for_online_recommendation_b200 as pkg
lib = pkg.require_cuda()
sizes = bench.feature_sizes('cfg5'); B = 8192
torch.manual_seed(0)
m = pkg.DeepFMAdam(sizes, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
host = bench.synth_batches(sizes, B, 3, 1)
enc = [m.encode(Xi, None, Y) for Xi, Y in host]
for i in range(4): m._fm_step(enc[i % 3], 0, enc[(i + 1) % 3])
torch.cuda.synchronize()
dbg = torch.zeros(8 + 4000 * 8, dtype=torch.int64, device='cuda')
lib.fmb_debug_set_runs_buffer.argtypes = [C.c_void_p]; lib.fmb_debug_set_runs_buffer.restype = None
lib.fmb_debug_set_runs_buffer(C.c_void_p(dbg.data_ptr()))
m._fm_step(enc[1], 0, enc[2]); torch.cuda.synchronize()
lib.fmb_debug_set_runs_buffer(None)
d = dbg.cpu().numpy(); n = int(d[0]); r = d[8:8 + min(n, 4000) * 8].reshape(-1, 8)
print('runs', n)
t0 = r[:, 1].min()
cyc = 1000 / 1965
end = r[:, 1] - t0 + (r[:, 2] + r[:, 3] + r[:, 4]) * cyc
print('first start 0 ns, last start %d ns, last end %d ns' % (r[:, 1].max() - t0, end.max()))
order = np.argsort(-end)
print('len start_ns end_ns direct ring update | c_issue c_wait c_cons (cycles)')
for i in order[:15]: print(r[i, 0], r[i, 1] - t0, int(end[i]), r[i, 2], r[i, 3], r[i, 4], '|', r[i, 5], r[i, 6], r[i, 7])
print('by length: lo hi count start_ns(mean) mean_direct mean_ring mean_update | issue wait cons')
for lo, hi in [(2, 3), (3, 8), (8, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 2000), (2000, 100000)]:
    s = (r[:, 0] >= lo) & (r[:, 0] < hi)
    if s.any(): print(lo, hi, int(s.sum()), int((r[s, 1] - t0).mean()), r[s, 2].mean().round(), r[s, 3].mean().round(), r[s, 4].mean().round(), '|', r[s, 5].mean().round(), r[s, 6].mean().round(), r[s, 7].mean().round())
