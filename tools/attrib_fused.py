"""Timing attribution of fm_step_fused_kernel: run it with phases switched off (debug hook) at the bench's cfg5 shape."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
from bench import feature_sizes, synth_batches
lib = pkg.require_cuda()
lib.fmb_debug_set_step_flags.argtypes = [C.c_int]
sizes = feature_sizes("cfg5"); F, k = len(sizes), 10
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
if len(sys.argv) > 2:
    with torch.no_grad(): m._table.mul_(float(sys.argv[2]))
enc = [m.encode(Xi, None, Y) for Xi, Y in synth_batches(sizes, B, 8, 1)]
p = lambda t: C.c_void_p(t.data_ptr())
N = B * F
sk = torch.empty(N, dtype=torch.int32, device="cuda"); pm = torch.empty_like(sk); pf = torch.empty_like(sk)
delta = torch.empty(B, device="cuda"); lossv = torch.empty(B, device="cuda")
bwsb = lib.fmb_bwd_workspace_bytes(N, k); bws = torch.empty(bwsb, dtype=torch.uint8, device="cuda")
wsb = lib.fmb_sort_workspace_bytes(N); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for flags in (0, 1, 2, 3, 0):
    lib.fmb_debug_set_step_flags(flags)
    tot = 0.0
    for i in range(12):
        e = enc[i % 8]
        if B <= 65536:
            lib.fmb_sort_fields(p(e.ids), B, F, p(m._field_off_dev), p(sk), p(pm), st)
        else:
            lib.fmb_sort_segment(p(e.ids), N, m._key_bits, p(ws), wsb, p(sk), p(pm), None, None, st)
        lib.fmb_pos_flags(p(sk), p(pm), N, p(pf), st)
        ev[0].record()
        rc = lib.fmb_fm_step_fused(p(e.ids), None, p(e.y), p(m._table), p(m.bias), p(pf), B, F, k, 0, m._lr, 0, p(delta),
                                   p(lossv), p(bws), bwsb, st)
        assert rc == 0, lib.fmb_last_error()
        ev[1].record()
        torch.cuda.synchronize()
        if i >= 2: tot += ev[0].elapsed_time(ev[1]) / 10
    print(f"flags {flags}: fused kernel {tot*1e3:.1f} us", flush=True)
lib.fmb_debug_set_step_flags(0)
