import sys; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import _lib
ctx=_lib.Context(0)
rng=np.random.default_rng(0)
for (M,N,K) in [(128,64,64),(200,256,256),(1000,768,256),(333,100,512),(5000,256,256),(10000,256,256),(5000,512,512),(5000,768,256)]:
    A=rng.standard_normal((M,K)).astype(np.float32); B=(rng.standard_normal((N,K))*0.06).astype(np.float32); bias=rng.standard_normal(N).astype(np.float32)
    ref=(A.astype(np.float64)@B.astype(np.float64).T+bias)
    for mode,name in ((2,'k_gemm_tc'),(3,'k_gemm_tma')):
        C=np.zeros((M,N),np.float32)
        for rep in range(2):
            ctx.profile_start(name)
            rc=ctx.lib.b2_debug_gemm_host(ctx.handle, mode, _lib.ptr(A), _lib.ptr(B), _lib.ptr(bias), _lib.ptr(C), M,N,K)
            ms,n,w=ctx.profile_stop()
        if rc!=0: print('mode',mode,'rc',rc, ctx.lib.b2_last_error(ctx.handle).decode()); continue
        print((M,N,K),name,'rel err', float(np.abs(C-ref).max()/np.abs(ref).max()), f'time {ms*1e3:.1f} us -> {2*M*N*K/ms/1e9:.1f} TF/s')
