import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
lib = pkg.require_cuda()
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
torch.manual_seed(0)
for (M, N, K) in [(128, 128, 32), (256, 128, 64), (300, 400, 400), (8192, 400, 10), (8192, 400, 400), (1000, 72, 8192), (400, 400, 8192), (77, 209, 133)]:
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda'); Cc = torch.zeros(M, N, device='cuda')
    rc = lib.fmb_gemm_tc_nt(P(A), P(B), P(Cc), M, N, K, None)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (Cc.double() - ref).abs().max().item(); rel = err / ref.abs().max().item()
    print('NT', (M, N, K), 'rc', rc, 'tc_error', lib.fmb_gemm_tc_error(), 'rel', rel)
# TN with colsum (dW): A(m=o,k=b) = gp[b][o], B(k=b,n=i) = x[b][i]
Bt, H = 8192, 400
gp = torch.randn(Bt, H, device='cuda'); x = torch.randn(Bt, H, device='cuda')
gW = torch.zeros(H, H, device='cuda'); gc = torch.zeros(H, device='cuda')
rc = lib.fmb_gemm_tc_strided(P(gp), 1, H, P(x), H, 1, P(gW), H, H, H, Bt, 0, None, None, 0, P(gc), None)
torch.cuda.synchronize()
ref = gp.double().t() @ x.double()
print('TN rc', rc, 'rel', ((gW.double() - ref).abs().max() / ref.abs().max()).item(), 'colsum rel',
      ((gc.double() - gp.double().sum(0)).abs().max() / gp.double().sum(0).abs().max()).item(), 'err', lib.fmb_gemm_tc_error())
# NN with mask (dX): A(m=b,k=o) = gp[b][o], B(k=o,n=i) = W[o][i]
W = torch.randn(H, H, device='cuda'); mask = torch.randn(Bt, H, device='cuda'); gx = torch.zeros(Bt, H, device='cuda')
rc = lib.fmb_gemm_tc_strided(P(gp), H, 1, P(W), H, 1, P(gx), H, Bt, H, H, 2, None, P(mask), H, None, None)
torch.cuda.synchronize()
ref = (gp.double() @ W.double()) * (mask > 0)
print('NN rc', rc, 'rel', ((gx.double() - ref).abs().max() / ref.abs().max()).item(), 'err', lib.fmb_gemm_tc_error())
# NT with bias+relu
bias = torch.randn(H, device='cuda'); y = torch.zeros(Bt, H, device='cuda')
rc = lib.fmb_gemm_tc_strided(P(x), H, 1, P(W), 1, H, P(y), H, Bt, H, H, 1, P(bias), None, 0, None, None)
torch.cuda.synchronize()
ref = torch.relu(x.double() @ W.double().t() + bias.double())
print('NT relu rc', rc, 'rel', ((y.double() - ref).abs().max() / ref.abs().max()).item(), 'err', lib.fmb_gemm_tc_error())

def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
fl = 2 * Bt * H * H
t = timeit(lambda: lib.fmb_gemm_tc_strided(P(x), H, 1, P(W), 1, H, P(y), H, Bt, H, H, 1, P(bias), None, 0, None, None))
print('fwd  NT  %.1f us  %.1f TFLOP/s' % (t, fl / t / 1e6))
t = timeit(lambda: lib.fmb_gemm_tc_strided(P(gp), H, 1, P(W), H, 1, P(gx), H, Bt, H, H, 2, None, P(mask), H, None, None))
print('dX   NN  %.1f us  %.1f TFLOP/s' % (t, fl / t / 1e6))
t = timeit(lambda: lib.fmb_gemm_tc_strided(P(gp), 1, H, P(x), H, 1, P(gW), H, H, H, Bt, 0, None, None, 0, P(gc), None))
print('dW   TN  %.1f us  %.1f TFLOP/s' % (t, fl / t / 1e6))
t = timeit(lambda: torch.mm(x, W.t()))
print('torch.mm fp32 (cuBLAS) %.1f us' % t)
torch.backends.cuda.matmul.allow_tf32 = True
t = timeit(lambda: torch.mm(x, W.t()))
print('torch.mm tf32 (cuBLAS) %.1f us' % t)
# through the MLP forward/backward (cfg4 tower)
B, k, L = 8192, 10, 3
bi = torch.randn(B, k, device='cuda'); n = lib.fmb_mlp_numel(k, L, H); mlp = (torch.rand(n, device='cuda') - 0.5) * 0.1
act = torch.empty(L, B, H, device='cuda'); head = torch.empty(L, B, device='cuda')
gmlp = torch.zeros(n, device='cuda'); gbi = torch.zeros(B, k, device='cuda'); gtop = torch.randn(B, device='cuda')
wsb = lib.fmb_mlp_bwd_workspace_bytes(B, H); ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
res = {}
lib.fmb_set_tensor_cores(0)
lib.fmb_mlp_forward(P(bi), k, P(mlp), B, k, L, H, P(act), P(head), None)
act0 = act.clone()
for tc in (0, 1):
    lib.fmb_set_tensor_cores(tc)
    lib.fmb_mlp_backward(P(bi), k, P(mlp), P(act0), P(gtop), L - 1, B, k, L, H, P(gmlp), P(gbi), k, P(ws), wsb, None)
    torch.cuda.synchronize()
    res[tc] = (gmlp.clone(), gbi.clone())
for i, nm in enumerate(('gmlp(same act)', 'gbi(same act)')):
    d = (res[0][i] - res[1][i]).abs().max().item(); s = res[0][i].abs().max().item()
    print(nm, 'max abs diff', d, 'rel', d / s)
for tc in (0, 1):
    lib.fmb_set_tensor_cores(tc)
    f = lambda: lib.fmb_mlp_forward(P(bi), k, P(mlp), B, k, L, H, P(act), P(head), None)
    b = lambda: lib.fmb_mlp_backward(P(bi), k, P(mlp), P(act), P(gtop), L - 1, B, k, L, H, P(gmlp), P(gbi), k, P(ws), wsb, None)
    tf = timeit(f, 10); tb = timeit(b, 10)
    res[tc] = (act.clone(), gmlp.clone(), gbi.clone())
    print('tc' if tc else 'simt', 'mlp fwd %.1f us  bwd %.1f us' % (tf, tb))
for i, nm in enumerate(('act', 'gmlp', 'gbi')):
    d = (res[0][i] - res[1][i]).abs().max().item(); s = res[0][i].abs().max().item()
    print(nm, 'max abs diff', d, 'rel', d / s)
print('tc_error', lib.fmb_gemm_tc_error())
