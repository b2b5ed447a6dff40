import sys, time; sys.path.insert(0,'.')
import numpy as np
from gtsfm_b200 import synthetic as syn, _lib
from gtsfm_b200.detector_descriptor import SuperPointEngine
ctx=_lib.Context(0)
sp = SuperPointEngine(syn.superpoint_state_dict(0), ctx=ctx)
for (H,W) in [(480,640),(1024,1024)]:
    img = syn.synthetic_frame(1,H,W)
    for i in range(3): xy,sc = sp.detect(img)
    t=time.perf_counter(); 
    for i in range(5): xy,sc = sp.detect(img)
    dt=(time.perf_counter()-t)/5
    print((H,W),'detect host wall ms', dt*1e3, 'kpts', len(xy))
    for k in ['k_conv_tma','k_conv3x3','k_conv1a','k_nms','k_head','k_compact','k_scan','k_to_gray']:
        ctx.profile_start(k); sp.detect(img); ms,n,w = ctx.profile_stop()
        if n: print(f'   {k:12s} {ms:7.3f} ms launches {n:3d} work {w/1e9:7.1f} GFLOP -> {w/1e9/ms if ms else 0:7.1f} TFLOP/s')
