"""Round-2 baseline capture of the kernels round 1 left unprofiled: SuperGlue (Sinkhorn) at 2048 keypoints, RANSAC-5pt at
K = 2000, the descriptor sampler and the LightGlue assignment (inside one 1024-keypoint pair).

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_misc.csv python profiles/capture_r02_misc.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gtsfm_b200 import _lib, synthetic as syn  # noqa: E402
from gtsfm_b200.detector_descriptor import SuperPointEngine  # noqa: E402
from gtsfm_b200.matcher import LightGlueEngine, SuperGlueEngine  # noqa: E402
from gtsfm_b200.verifier import RansacEngine  # noqa: E402
from oracle import verifier_ref  # noqa: E402  (scene generator only)

ctx = _lib.Context(0)
sg = SuperGlueEngine(syn.superglue_state_dict(1, "sharp"), ctx=ctx)
kp0, sc0, d0, kp1, sc1, d1, _ = syn.synthetic_features(12, 2048, 1900)
for _ in range(2):
    m = sg.match(kp0, sc0, d0, kp1, sc1, d1, (480, 640, 3), (480, 640, 3))
print("superglue matches", len(m))
lg = LightGlueEngine(syn.lightglue_state_dict(2, "bench"), ctx=ctx)
a = syn.synthetic_features(12, 1024, 1024)
print("lightglue matches", len(lg.match(a[0], a[2], a[3], a[5])))
rs = RansacEngine(ctx=ctx)
k1, k2, matches, K, R, t, is_in = verifier_ref.synthetic_two_view(3, 2000, 0.5)
n1 = verifier_ref.calibrate(k1[matches[:, 0]], *K)
n2 = verifier_ref.calibrate(k2[matches[:, 1]], *K)
for _ in range(2):
    E, mask, Rr, tr = rs.essential(n1, n2, 4.0 / K[0])
print("ransac inliers", int(mask.sum()))
sp = SuperPointEngine(syn.superpoint_state_dict(0), ctx=ctx)
frames, _ = syn.synthetic_sequence(2, 480, 640)
xy, sc = sp.detect(frames[0])
sp.describe(xy[:5000])
print("launches", ctx.launch_count())
