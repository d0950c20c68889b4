"""Host-side trace of the e2e loop of bench.py (fmb_session_fm_step_host_async, four slots): time of every submit / collect call."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
from bench import feature_sizes, synth_batches
lib = pkg.require_cuda()
sizes = feature_sizes("cfg5"); F, k, B, NB = len(sizes), 10, 8192, 16
torch.manual_seed(0)
model = pkg.DeepFMAdam(sizes, embedding_size=k, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
host = synth_batches(sizes, B, NB, 1234)
host_ids = [np.ascontiguousarray((Xi + model._offsets_np[:-1][None, :]).astype(np.int32)) for Xi, _ in host]
pin_ids = [torch.from_numpy(a).pin_memory() for a in host_ids]
pin_y = [torch.from_numpy(np.ascontiguousarray(y)).pin_memory() for _, y in host]
sess = model._get_session(B)
tptr, bptr = C.c_void_p(model._table.data_ptr()), C.c_void_p(model.bias.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
loss = C.c_float()
NSLOT = lib.fmb_session_host_slots()
def submit(i):
    j = i % NB
    assert lib.fmb_session_fm_step_host_async(sess, i % NSLOT, C.c_void_p(pin_ids[j].data_ptr()), None, C.c_void_p(pin_y[j].data_ptr()),
                                              B, tptr, bptr, model._key_bits, 0, model._lr, 0, st) == 0
def collect(i):
    assert lib.fmb_session_wait_loss(sess, i % NSLOT, C.byref(loss)) == 0
def run_host(n, base, trace=None):
    for i in range(n):
        t0 = time.perf_counter(); submit(base + i); t1 = time.perf_counter()
        if i >= NSLOT - 1: collect(base + i - (NSLOT - 1))
        t2 = time.perf_counter()
        if trace is not None: trace.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6))
    for i in range(max(0, n - (NSLOT - 1)), n): collect(base + i)
run_host(12, 0); torch.cuda.synchronize()
for K in (20, 20, 200):
    tr = []
    t0 = time.perf_counter(); run_host(K, 1000, tr); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("K=%d: %.1f us/step (%.1f M samples/s); submit us: first 6 %s median %.1f; collect us: first 6 %s median %.1f" % (
        K, dt / K * 1e6, B * K / dt / 1e6, [round(a) for a, _ in tr[:6]], np.median([a for a, _ in tr]), [round(b) for _, b in tr[:6]], np.median([b for _, b in tr])))
