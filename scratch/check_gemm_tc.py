import sys; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import _lib
ctx=_lib.Context(0)
rng=np.random.default_rng(0)
for (M,N,K) in [(128,128,64),(200,256,256),(1000,768,256),(333,100,512)]:
    A=rng.standard_normal((M,K)).astype(np.float32); B=(rng.standard_normal((N,K))*0.06).astype(np.float32); bias=rng.standard_normal(N).astype(np.float32)
    ref=(A.astype(np.float64)@B.astype(np.float64).T+bias).astype(np.float64)
    outs={}
    for mode in (0,1,2):
        C=np.zeros((M,N),np.float32)
        rc=ctx.lib.b2_debug_gemm_host(ctx.handle, mode, _lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M,N,K)
        if rc!=0: print('mode',mode,'rc',rc, ctx.lib.b2_last_error(ctx.handle).decode()); continue
        outs[mode]=C
        print((M,N,K),'mode',mode,'max abs err vs fp64', float(np.abs(C-ref).max()), 'rel', float(np.abs(C-ref).max()/np.abs(ref).max()))
