import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from gtsfm_b200 import synthetic as syn, _lib
from gtsfm_b200.matcher import LightGlueEngine
ctx=_lib.Context(0)
lg = LightGlueEngine(syn.lightglue_state_dict(2,'bench'), ctx=ctx)
N=int(sys.argv[1]) if len(sys.argv)>1 else 5000
kp0,sc0,d0,kp1,sc1,d1,gt = syn.synthetic_features(3,N,N)
for i in range(2): lg.match(kp0,d0,kp1,d1)
t=time.perf_counter(); lg.match(kp0,d0,kp1,d1); print('host wall per pair ms', (time.perf_counter()-t)*1e3)
tot=0
for k in ['k_flash','k_gemm','k_lg_col_argmax','k_lg_col_stats','k_lg_row','k_lg_ln_gelu','k_lg_split','k_lg_rowheads','k_lg_prune','k_lg_gather','k_lg_posenc','k_lg_filter']:
    ctx.profile_start(k); lg.match(kp0,d0,kp1,d1); ms,n,w = ctx.profile_stop(); tot+=ms
    print(f'{k:18s} {ms:8.3f} ms  launches {n:4d}  work {w/1e9:9.1f} GFLOP  -> {w/1e9/ms if ms else 0:8.1f} TFLOP/s')
print('sum', tot)
