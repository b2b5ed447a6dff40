"""Per-kernel device time of one 5-pt verification (900 matches, 1000 hypotheses): python scratch/prof_ransac.py"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gtsfm_b200 import _lib
from gtsfm_b200.verifier import RansacEngine
from oracle import verifier_ref as vr

kp1, kp2, m, K, *_ = vr.synthetic_two_view(5, 900, 0.9)
n1, n2 = vr.calibrate(kp1[m[:, 0]], *K), vr.calibrate(kp2[m[:, 1]], *K)
ctx = _lib.Context(0)
eng = RansacEngine(ctx=ctx)
for _ in range(3):
    eng.essential(n1, n2, 4.0 / K[0])
for pre in ("k_rs_hyp", "k_rs_score", "k_rs_refine", "k_rs_pose", "k_rs"):
    ctx.profile_start(pre)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.essential(n1, n2, 4.0 / K[0])
    wall = (time.perf_counter() - t0) / 5
    ms, n, _ = ctx.profile_stop()
    print(f"{pre:12s} {ms / 5:8.3f} ms / call  ({n // 5} launches)   host wall {wall * 1e3:.3f} ms")
