"""Time the fp32 gate+residual epilogue GEMM (in place, bf16 copy) vs the plain bf16 epilogue for each tile config."""
import sys; sys.path.insert(0,'.')
import torch
from instarevive_b200 import _lib
L=_lib.lib(); P=_lib.ptr; S=_lib.stream_ptr; dev='cuda'
def t(fn,it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(True),torch.cuda.Event(True); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for (M,N,K,T) in [(4096,1152,1152,4096),(4096,1152,4608,4096),(25600,1152,1152,1024),(25600,1152,4608,1024)]:
    A=torch.randn(M,K,device=dev).bfloat16(); W=(torch.randn(N,K,device=dev)*0.02).bfloat16(); b=torch.randn(N,device=dev)
    o=torch.empty(M,N,device=dev,dtype=torch.bfloat16); x=torch.randn(M,N,device=dev); gate=torch.randn(M//T,6*N,device=dev)
    for cfg in (128,256,2128,2256):
        ms0=t(lambda: L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,0,1.0,P(o),None,None,None,0,1,cfg,S()))
        ms2=t(lambda: L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,2,1.0,P(o),P(x),P(x),gate.data_ptr()+2*N*4,6*N,T,cfg,S()))
        ms2n=t(lambda: L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,2,1.0,None,P(x),P(x),None,0,1,cfg,S()))
        f=2.0*M*N*K/1e9
        print(f"M{M} N{N} K{K} cfg{cfg}: bf16 {ms0*1e3:7.1f} us {f/ms0:6.0f} TF | f32+gate+resid+copy {ms2*1e3:7.1f} us {f/ms2:6.0f} TF | f32+resid {ms2n*1e3:7.1f} us {f/ms2n:6.0f} TF")
