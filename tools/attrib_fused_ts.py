"""Per-CTA phase timestamps of fm_step_fused_kernel (debug hook): where does a tile's time go?"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
from bench import feature_sizes, synth_batches
lib = pkg.require_cuda()
sizes = feature_sizes("cfg5"); F, k = len(sizes), 10
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
enc = [m.encode(Xi, None, Y) for Xi, Y in synth_batches(sizes, B, 4, 1)]
p = lambda t: C.c_void_p(t.data_ptr())
N = B * F
sk = torch.empty(N, dtype=torch.int32, device="cuda"); pm = torch.empty_like(sk); pf = torch.empty_like(sk)
delta = torch.empty(B, device="cuda"); lossv = torch.empty(B, device="cuda")
bwsb = lib.fmb_bwd_workspace_bytes(N, k); bws = torch.empty(bwsb, dtype=torch.uint8, device="cuda")
grid = (B + 3) // 4        # upper bound on the number of tiles (>= 4 samples each)
ts = torch.zeros(grid * 8 + 64, dtype=torch.int64, device="cuda")
lib.fmb_debug_set_step_timestamps.argtypes = [C.c_void_p]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    e = enc[i % 4]
    lib.fmb_sort_fields(p(e.ids), B, F, p(m._field_off_dev), p(sk), p(pm), None)
    lib.fmb_pos_flags(p(sk), p(pm), N, p(pf), None)
    if i == 5: lib.fmb_debug_set_step_timestamps(p(ts))
    torch.cuda.synchronize()
    ev0.record()
    rc = lib.fmb_fm_step_fused(p(e.ids), None, p(e.y), p(m._table), p(m.bias), p(pf), B, F, k, 0, m._lr, 0, p(delta),
                               p(lossv), p(bws), bwsb, None)
    assert rc == 0
    ev1.record()
    torch.cuda.synchronize()
    print("step", i, "fused kernel by CUDA events: %.1f us" % (ev0.elapsed_time(ev1) * 1e3))
lib.fmb_debug_set_step_timestamps(None)
t = ts.cpu().numpy()[:grid * 8].reshape(grid, 8)[:, :7].astype(np.float64)
t = t[t[:, 0] > 0]
grid = len(t)
t0 = t[:, 0].min()
names = ["ids+pos+compaction", "gather", "reduce S,Q", "logit+loss", "single-hit updates", "multi staging"]
print("CTAs", grid, "start spread (us): min 0, median %.2f, max %.2f" % ((np.median(t[:, 0]) - t0) / 1e3, (t[:, 0].max() - t0) / 1e3))
print("end (us): median %.2f max %.2f" % ((np.median(t[:, 6]) - t0) / 1e3, (t[:, 6].max() - t0) / 1e3))
for i, n in enumerate(names):
    d = (t[:, i + 1] - t[:, i]) / 1e3
    print("%-22s median %.2f  p90 %.2f  max %.2f us" % (n, np.median(d), np.percentile(d, 90), d.max()))
