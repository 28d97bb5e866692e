import sys; sys.path.insert(0,'.')
import torch
from instarevive_b200 import _lib
L=_lib.lib(); P=_lib.ptr; S=_lib.stream_ptr; dev='cuda'
def t(fn,it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(True),torch.cuda.Event(True); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for (M,N,K) in [(4096,1152,1152),(4096,4608,1152),(4096,1152,4608),(25600,1152,1152),(25600,4608,1152)]:
    A=torch.randn(M,K,device=dev).bfloat16(); W=torch.randn(N,K,device=dev).bfloat16(); b=torch.randn(N,device=dev); o=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    for bn in (64,128,256):
        ms=t(lambda: L.ir_gemm_bf16(P(A),P(W),P(b),M,N,K,1,0,0,0,0,1.0,P(o),None,None,None,0,1,bn,S()))
        tiles=((M+127)//128)*((N+bn-1)//bn); byts=tiles*((K+63)//64)*(16384+bn*128)
        print(f"M{M} N{N} K{K} bn{bn}: {ms*1e3:.1f} us  L2->SM {byts/ms/1e9:.2f} TB/s  ({2.0*M*N*K/ms/1e9:.0f} TFLOP/s equivalent)")
