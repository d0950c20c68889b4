import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_for_online_recommendation_b200 as pkg
lib=pkg.require_cuda()
lib.fmb_gemm_tc_nt.argtypes=[C.c_void_p]*3+[C.c_int]*3+[C.c_void_p]; lib.fmb_gemm_tc_nt.restype=C.c_int
lib.fmb_gemm_tc_error.restype=C.c_int
torch.manual_seed(0)
for (M,N,K) in [(128,128,32),(256,128,64),(300,400,400),(8192,400,10),(8192,400,400)]:
    A=torch.randn(M,K,device='cuda'); B=torch.randn(N,K,device='cuda'); Cc=torch.zeros(M,N,device='cuda')
    rc=lib.fmb_gemm_tc_nt(C.c_void_p(A.data_ptr()),C.c_void_p(B.data_ptr()),C.c_void_p(Cc.data_ptr()),M,N,K,None)
    torch.cuda.synchronize()
    ref=(A.double()@B.double().t()).float()
    err=(Cc-ref).abs().max().item(); rel=err/ref.abs().max().item()
    print((M,N,K),'rc',rc,'tc_error',lib.fmb_gemm_tc_error(),'max abs err',err,'rel',rel, 'nonzero', int((Cc!=0).sum()))
# timing vs SIMT through the MLP forward (cfg4 tower)
lib.fmb_set_tensor_cores.argtypes=[C.c_int]
B,k,L,H=8192,10,3,400
bi=torch.randn(B,k,device='cuda'); n=H*k+H+(L-1)*(H*H+H); mlp=(torch.rand(n,device='cuda')-0.5)*0.1
act=torch.empty(L,B,H,device='cuda'); head=torch.empty(L,B,device='cuda')
lib.fmb_mlp_forward.argtypes=[C.c_void_p,C.c_int,C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.c_void_p,C.c_void_p,C.c_void_p]
for tc in (0,1):
    lib.fmb_set_tensor_cores(tc)
    for _ in range(3): lib.fmb_mlp_forward(C.c_void_p(bi.data_ptr()),k,C.c_void_p(mlp.data_ptr()),B,k,L,H,C.c_void_p(act.data_ptr()),C.c_void_p(head.data_ptr()),None)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lib.fmb_mlp_forward(C.c_void_p(bi.data_ptr()),k,C.c_void_p(mlp.data_ptr()),B,k,L,H,C.c_void_p(act.data_ptr()),C.c_void_p(head.data_ptr()),None)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    fl=2*B*(k*H+(L-1)*H*H)
    print('tc' if tc else 'simt','mlp fwd',round(ms*1000,1),'us',round(fl/ms/1e9,2),'TFLOP/s (fp32-equivalent)', float(act.abs().sum()))
