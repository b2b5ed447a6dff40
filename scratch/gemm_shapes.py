import sys; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import _lib
ctx=_lib.Context(0)
rng=np.random.default_rng(0)
for (M,N,K) in [(128,64,64),(128,64,256),(128,64,512),(128,256,256),(5000,256,256),(10000,256,256),(5000,256,512),(5000,512,512),(5000,768,256),(18944,256,256),(18944,64,256)]:
    A=rng.standard_normal((M,K)).astype(np.float32); B=(rng.standard_normal((N,K))*0.06).astype(np.float32)
    C=np.zeros((M,N),np.float32)
    for rep in range(2):
        ctx.profile_start('k_gemm_tc')
        rc=ctx.lib.b2_debug_gemm_host(ctx.handle, 1, _lib.ptr(A), _lib.ptr(B), None, _lib.ptr(C), M,N,K)
        ms,n,w=ctx.profile_stop()
    ctas=((M+127)//128)*((N+63)//64)
    print(f'M={M:6d} N={N:4d} K={K:4d} ctas={ctas:5d} time={ms*1e3:8.1f} us  -> {2*M*N*K/ms/1e9:7.1f} TFLOP/s')
