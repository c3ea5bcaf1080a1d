import sys, os, ctypes as C
sys.path[:0]=['/root/repo','/root/repo/test-time-adaptation-asr-suta_b200']
import torch
from suta_b200 import _lib
from suta_b200._lib import check
lib=_lib.load()
def P(t): return C.c_void_p(t.data_ptr()) if t is not None else None
M,N,K=20096,4096,4096
a=torch.randn(M,K,device='cuda').bfloat16(); b=torch.randn(N,K,device='cuda').bfloat16(); o=torch.empty(M,N,device='cuda',dtype=torch.bfloat16)
st=torch.cuda.current_stream().cuda_stream
for i in range(2):
    check(lib.suta_op_gemm(P(a),M,K,P(b),N,K,M,N,K,None,P(o),N,None,None,N,0,None,None,N,st))
torch.cuda.synchronize()
