import sys; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import _lib
ctx=_lib.Context(0)
rng=np.random.default_rng(0)
for (M,N,K) in [(128,64,64),(128,64,256),(5000,256,256)]:
    A=rng.standard_normal((M,K)).astype(np.float32); B=(rng.standard_normal((N,K))*0.06).astype(np.float32)
    C=np.zeros((M,N),np.float32)
    for rep in range(2):
        print((M,N,K), flush=True)
        rc=ctx.lib.b2_debug_gemm_host(ctx.handle, 1, _lib.ptr(A), _lib.ptr(B), None, _lib.ptr(C), M,N,K)
