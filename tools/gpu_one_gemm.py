import sys; sys.path.insert(0,'.')
import torch
from instarevive_b200 import _lib
L=_lib.lib(); P=_lib.ptr; S=_lib.stream_ptr; dev='cuda'
M,N,K,T=25600,1152,1152,1024
cfg=int(sys.argv[1]) if len(sys.argv)>1 else 2256
A=torch.randn(M,K,device=dev).bfloat16(); W=(torch.randn(N,K,device=dev)*0.02).bfloat16(); b=torch.randn(N,device=dev)
o=torch.empty(M,N,device=dev,dtype=torch.bfloat16); x=torch.randn(M,N,device=dev); gate=torch.randn(M//T,6*N,device=dev)
for _ in range(3):
    L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,2,1.0,P(o),P(x),P(x),gate.data_ptr()+2*N*4,6*N,T,cfg,S())
torch.cuda.synchronize()
torch.cuda.profiler.start()
L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,2,1.0,P(o),P(x),P(x),gate.data_ptr()+2*N*4,6*N,T,cfg,S())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
