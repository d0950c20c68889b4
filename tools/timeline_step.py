"""Timeline of ONE graph-replayed FM step (cfg5, B = 8192): when does each kernel start and end (globaltimer ns)?
The debug hooks are set before the first step so that they are baked into the captured graph."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
from bench import feature_sizes, synth_batches
lib = pkg.require_cuda()
sizes = feature_sizes("cfg5"); F, k = len(sizes), 10
B = 8192
torch.manual_seed(0)
m = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
enc = [m.encode(Xi, None, Y) for Xi, Y in synth_batches(sizes, B, 4, 1)]
grid = (B + 3) // 4
ts = torch.zeros(grid * 8 + 64, dtype=torch.int64, device="cuda")
rd = torch.zeros(8 + 4000 * 8, dtype=torch.int64, device="cuda")
sd = torch.zeros(16, dtype=torch.int64, device="cuda")
for fn in ("fmb_debug_set_step_timestamps", "fmb_debug_set_runs_buffer", "fmb_debug_set_sort_buffer"):
    getattr(lib, fn).argtypes = [C.c_void_p]; getattr(lib, fn).restype = None
lib.fmb_debug_set_step_timestamps(C.c_void_p(ts.data_ptr()))
lib.fmb_debug_set_runs_buffer(C.c_void_p(rd.data_ptr()))
lib.fmb_debug_set_sort_buffer(C.c_void_p(sd.data_ptr()))
BIG = (1 << 62)
def reset():
    ts.zero_(); rd.zero_(); sd.zero_(); sd[14] = BIG; sd[13] = BIG
for i in range(40):
    m._fm_step(enc[i % 4], 0, enc[(i + 1) % 4])
torch.cuda.synchronize()
for rep in range(3):
    reset(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    m._fm_step(enc[(40 + 2 * rep) % 4], 0, enc[(41 + 2 * rep) % 4])
    ev[1].record()
    m._fm_step(enc[(41 + 2 * rep) % 4], 0, enc[(42 + 2 * rep) % 4])      # second step: stamps are overwritten by it
    ev[2].record()
    torch.cuda.synchronize()
    t = ts.cpu().numpy()[:grid * 8].reshape(grid, 8)[:, :7]
    t = t[t[:, 0] > 0]
    r = rd.cpu().numpy(); n = int(r[0]); rr = r[8:8 + min(n, 4000) * 8].reshape(-1, 8)
    s = sd.cpu().numpy()
    # the runs records of BOTH steps are in rr; keep those of the second step (start after the fused kernel's first CTA)
    f0, f1 = t[:, 0].min(), t[:, 6].max()
    cyc = 1000 / 1965
    rs = rr[:, 1]; re = rr[:, 1] + (rr[:, 2] + rr[:, 3] + rr[:, 4]) * cyc
    sel = rs > f0
    print("events: step A %.1f us, step B %.1f us" % (ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3))
    print("step B timeline (us, 0 = first fused CTA starts):")
    print("  fused: last CTA start %.1f, first end %.1f, median end %.1f, last end %.1f  (%d CTAs)" % (
        (t[:, 0].max() - f0) / 1e3, (t[:, 6].min() - f0) / 1e3, (np.median(t[:, 6]) - f0) / 1e3, (f1 - f0) / 1e3, len(t)))
    if sel.any():
        print("  runs : first start %.1f, last start %.1f, last end %.1f  (%d runs)" % (
            (rs[sel].min() - f0) / 1e3, (rs[sel].max() - f0) / 1e3, (re[sel].max() - f0) / 1e3, int(sel.sum())))
    print("  sort (both steps: min start / max end): start %.1f, last small field done %.1f, first large field done %.1f, end %.1f" % (
        (s[14] - f0) / 1e3, (s[12] - f0) / 1e3, (s[13] - f0) / 1e3, (s[15] - f0) / 1e3))
