"""Timings of the other BASELINE.json configs on one B200 (CUDA events / wall clock for the persistent online kernels).
Prints one JSON object.  python tools/bench_configs.py"""
import sys, os, io, json, time, contextlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fm_for_online_recommendation_b200 as pkg
from bench import CRITEO_TINY
pkg.require_cuda()
FRAPPE = [957, 4082, 7, 7, 2, 3, 2, 9, 80, 233]
out = {}

def batches(sizes, B, nb, seed):
    rng = np.random.RandomState(seed)
    return [(np.stack([rng.randint(0, fs, size=B) for fs in sizes], 1).astype(np.int64),
             (rng.uniform(size=B) < 0.3).astype(np.float32)) for _ in range(nb)]

def time_steps(fn, n=200, warm=20):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000

# cfg1 shape: FM k=10 on ml-100k-shaped ids (943 users x 1682 items), mini-batch 256 (device-resident encoded batches)
torch.manual_seed(0)
m = pkg.FMAdam([943, 1682], embedding_size=10, n=0.01)
enc = [m.encode(Xi, None, Y) for Xi, Y in batches([943, 1682], 256, 8, 1)]
us = time_steps(lambda i: m._fm_step(enc[i % 8], 0))
out["cfg1_FMAdam_update_embedding_B256"] = {"us_per_step": round(us, 1), "samples_per_s": round(256 / us * 1e6)}
# cfg3 shape: NFM k=64, Frappe-shaped fields, batch 4096, tower 64-64 (one hidden layer of 64)
torch.manual_seed(0)
m = pkg.NFMAdam(FRAPPE, embedding_size=64, num_hidden_layers=1, neuron_per_hidden_layer=64, n=1e-3)
enc = [m.encode(Xi, None, Y) for Xi, Y in batches(FRAPPE, 4096, 8, 2)]
us = time_steps(lambda i: m._deep_fit(enc[i % 8]), n=100)
out["cfg3_NFMAdam_fit_B4096_k64"] = {"us_per_step": round(us, 1), "samples_per_s": round(4096 / us * 1e6)}
us = time_steps(lambda i: m._fm_step(enc[i % 8], 0), n=100)
out["cfg3_NFMAdam_update_embedding_B4096_k64"] = {"us_per_step": round(us, 1), "samples_per_s": round(4096 / us * 1e6)}
# cfg4 shape, FM-only step and hedge fit (DeepFMOnn, L=3, H=400, B=8192)
torch.manual_seed(0)
m = pkg.DeepFMAdam(CRITEO_TINY, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4)
enc = [m.encode(Xi, None, Y) for Xi, Y in batches(CRITEO_TINY, 8192, 8, 3)]
us = time_steps(lambda i: m._fm_step(enc[i % 8], 0), n=100)
out["cfg4_DeepFMAdam_update_embedding_B8192"] = {"us_per_step": round(us, 1), "samples_per_s": round(8192 / us * 1e6)}
us = time_steps(lambda i: m._deep_fit(enc[i % 8]), n=50, warm=5)
out["cfg4_DeepFMAdam_fit_B8192_tower400"] = {"us_per_step": round(us, 1), "samples_per_s": round(8192 / us * 1e6)}
torch.manual_seed(0)
m = pkg.DeepFMOnn(CRITEO_TINY, embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400, n=1e-4, batch_size=8192)
enc = [m.encode(Xi, None, Y) for Xi, Y in batches(CRITEO_TINY, 8192, 4, 4)]
us = time_steps(lambda i: m._hedge_fit(enc[i % 4]), n=20, warm=3)
out["cfg4_DeepFMOnn_hedge_fit_B8192_tower400"] = {"us_per_step": round(us, 1), "samples_per_s": round(8192 / us * 1e6)}
# online per-example mode (persistent kernel): the reference's scripts' shape (L=5, H=10, 2500 samples per call)
for name in ("FMAdam", "DeepFMAdam", "NFMOnn"):
    torch.manual_seed(0)
    kw = {} if name == "FMAdam" else dict(num_hidden_layers=5, neuron_per_hidden_layer=10)
    if name.endswith("Onn"): kw["batch_size"] = 1
    m = getattr(pkg, name)(CRITEO_TINY, embedding_size=10, n=1e-4, **kw)
    Xi, Y = batches(CRITEO_TINY, 2500, 1, 5)[0]
    Xv = np.ones_like(Xi, dtype=np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        m.run_experiment(Xi[:100], Xv[:100], Y[:100])   # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.run_experiment(Xi, Xv, Y)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"online_{name}_run_experiment_2500"] = {"seconds": round(dt, 4), "samples_per_s": round(2500 / dt)}
# cfg2 shape: classical fp64 learners, d = 8 (+ bias column), N = 59 535, m = 40
rng = np.random.RandomState(0)
N, d = 59535, 9
X = rng.standard_normal((N, d)); X[:, -1] = 1.0
y = np.sign(rng.standard_normal(N))
for name, eta in (("FM_FTRL", 0.005), ("SFTRL_CCFM", 0.005), ("SFTRL_Vanila", 0.005)):
    torch.manual_seed(0)
    mdl = getattr(pkg, name)(torch.DoubleTensor(X), torch.DoubleTensor(y), "cls", eta, 40)
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mdl.online_learning()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"cfg2_{name}_online_learning_N59535_m40"] = {"seconds": round(dt, 3), "samples_per_s": round(N / dt)}
print(json.dumps(out))
