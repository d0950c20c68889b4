"""Per-field timeline of the sort kernels of one eager FM-step sort (cfg5, B = 8192): globaltimer stamps per field."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
from fm_for_online_recommendation_b200._lib import RunList
from bench import feature_sizes, synth_batches
lib = pkg.require_cuda()
sizes = feature_sizes("cfg5"); F, k = len(sizes), 10
B = 8192; N = B * F
m = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
enc = [m.encode(Xi, None, Y) for Xi, Y in synth_batches(sizes, B, 4, 1)]
p = lambda t: C.c_void_p(t.data_ptr())
sk = torch.empty(N, dtype=torch.int32, device="cuda"); pm = torch.empty_like(sk); pf = torch.empty_like(sk)
nseg, cap = C.c_int(), C.c_int(); lib.fmb_runlist_shape(B, F, C.byref(nseg), C.byref(cap))
rl = torch.empty((nseg.value * cap.value, 4), dtype=torch.int32, device="cuda"); rc = torch.zeros(2 * nseg.value, dtype=torch.int32, device="cuda")
rld = RunList(rl.data_ptr(), rc.data_ptr(), nseg.value, cap.value)
sd = torch.zeros(16 + 10 * F, dtype=torch.int64, device="cuda")
lib.fmb_debug_set_sort_buffer.argtypes = [C.c_void_p]; lib.fmb_debug_set_sort_buffer.restype = None
for flags in (0, 1):
    for i in range(5):
        e = enc[i % 4]
        if i == 4:
            sd.zero_(); sd[14] = 1 << 62; sd[13] = 1 << 62
            lib.fmb_debug_set_sort_buffer(p(sd))
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        assert lib.fmb_sort_fields_ex(p(e.ids), B, F, p(m._field_off_dev), p(sk), p(pm), p(pf), C.byref(rld), flags, None) == 0
        ev1.record(); torch.cuda.synchronize()
    lib.fmb_debug_set_sort_buffer(None)
    s = sd.cpu().numpy()
    t0 = s[14]
    print("flags", flags, "events %.1f us; launch span %.1f us" % (ev0.elapsed_time(ev1) * 1e3, (s[15] - t0) / 1e3))
    for f in range(F):
        a, b, c = s[16 + 2 * f], s[17 + 2 * f], s[16 + 4 * F + f]
        sa, sb, sc = s[16 + 2 * F + 2 * f], s[17 + 2 * F + 2 * f], s[16 + 5 * F + f]
        line = "  field %2d rows %8d:" % (f, sizes[f])
        if a: line += " radix start %.1f passes done %.1f end %.1f" % ((a - t0) / 1e3, (c - t0) / 1e3, (b - t0) / 1e3)
        if sa: line += " sparse start %.1f loaded %.1f inserted %.1f compacted %.1f ranked %.1f end %.1f" % ((sa - t0) / 1e3, (s[16 + 6 * F + f] - t0) / 1e3, (s[16 + 7 * F + f] - t0) / 1e3, (sc - t0) / 1e3, (s[16 + 8 * F + f] - t0) / 1e3, (sb - t0) / 1e3)
        if f < 14 or f > 36: print(line)
