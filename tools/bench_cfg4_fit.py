"""BASELINE configs[3]: one DeepFMAdam.fit (fwd + tower bwd + table/tower/bias updates) at B = 8192, k = 10,
400-400-400 tower, 1 006 628-row tables; tensor-core tower vs the exact SIMT tower.  Prints one JSON line.
    python tools/bench_cfg4_fit.py [--steps 50] [--tc-only]"""
import sys, os, json, argparse, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fm_for_online_recommendation_b200 as pkg
from bench import CRITEO_TINY, synth_batches

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--tc-only", action="store_true")
args = ap.parse_args()
lib = pkg.require_cuda()
B, k, L, H = 8192, 10, 3, 400
NB = 8
host = synth_batches(CRITEO_TINY, B, NB, 7)
out = {"workload": "cfg4: DeepFMAdam.fit, B=8192, k=10, tower 10-400-400-400, F=39, 1006628 rows", "steps": args.steps}
for tc in ((1,) if args.tc_only else (1, 0)):
    lib.fmb_set_tensor_cores(tc)
    torch.manual_seed(0)
    m = pkg.DeepFMAdam(CRITEO_TINY, embedding_size=k, num_hidden_layers=L, neuron_per_hidden_layer=H, n=1e-4)
    enc = [m.encode(Xi, None, Y) for Xi, Y in host]
    for i in range(5):
        m._deep_fit(enc[i % NB])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        m._deep_fit(enc[i % NB])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.steps * 1000
    key = "tensor_core" if tc else "simt_exact"
    flop = 2 * B * (k * H + (L - 1) * H * H) * 3   # fwd + dX + dW
    out[key] = {"us_per_fit": us, "samples_per_s": B / us * 1e6, "tower_gflop_per_fit": flop / 1e9}
lib.fmb_set_tensor_cores(1)
out["tc_error"] = lib.fmb_gemm_tc_error()
print(json.dumps(out))
